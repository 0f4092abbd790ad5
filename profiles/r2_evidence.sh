#!/bin/bash
# round-2 evidence: full GPU suite, launch lists with DRAM bytes, full captures
# (exported to text on the box), default bench at N=1
cd /root/repo
O=gpurun_out/ev; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_gpu_tests.log
for w in cfg2 cfg3 cfg4 cfg5; do
  extra=""; [ $w = cfg4 ] && extra="--ranks phylum,genus,species --mode above --samples 8"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"wk::|fill_slots|compact|unpack" -c 60 --csv --log-file $O/r2_launches_$w.csv \
    python bench.py --workload $w $extra --steps 2 --warmup 1 --no-cpu --no-e2e > $O/launch_$w.log 2>&1
done
cap() { # kernel-regex name bench-args
  k=$1; n=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/$n python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > $O/cap_$n.log 2>&1
  ncu -i /tmp/$n.ncu-rep --page details > $O/r2_${n}_details.txt 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page raw --csv > $O/r2_${n}_raw.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page source --csv > $O/r2_${n}_source.csv 2>/dev/null
}
cap classify_seg_kernel seg_cfg2 --workload cfg2
cap classify_strata_kernel strata_cfg5 --workload cfg5
cap strata_apply_kernel apply_cfg5 --workload cfg5
cap ordinal_match_kernel ordinal_cfg3 --workload cfg3
cap classify_multi_kernel multi_cfg4 --workload cfg4 --ranks phylum,genus,species --mode above --samples 8
timeout 900 python bench.py > $O/r2_bench_default_1gpu.json 2> $O/bench_default.err
tail -3 $O/r2_gpu_tests.log; ls -la $O | head -40; cut -c1-600 $O/r2_bench_default_1gpu.json
