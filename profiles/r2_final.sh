#!/bin/bash
# end of round 2: smoke(), the GPU suite, the default bench line at N=1, the
# reference arm, the file bench, fresh captures of the cfg5 kernels
cd "$(dirname "$0")/.."
O=gpurun_out/fin; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/r2_gpu_tests.log
timeout 900 python bench.py > $O/r2_bench_default_1gpu.json 2> $O/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_1gpu.json 2> $O/bench_reference.err
timeout 900 python scratch/file_bench.py > $O/r2_file_bench.json 2> $O/file_bench.err
cap() { k=$1; n=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/$n python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > $O/cap_$n.log 2>&1
  ncu -i /tmp/$n.ncu-rep --page details > $O/r2_${n}_details.txt 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page raw --csv > $O/r2_${n}_raw.csv 2>/dev/null
}
cap classify_strata_kernel strata_cfg5 --workload cfg5
cap strata_apply_kernel apply_cfg5 --workload cfg5
tail -2 $O/smoke.log $O/r2_gpu_tests.log; cut -c1-300 $O/r2_bench_reference_1gpu.json; tail -2 $O/bench_reference.err
