#!/bin/bash
cd /root/repo
for o in "" "--opt strata_bpp=1184" "--opt strata_bpp=296"; do
  echo "== $o" >> gpurun_out/c26_cfg5.log
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" >> gpurun_out/c26_cfg5.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"strata|fill_slots" -c 6 --csv --log-file gpurun_out/c26_launches.csv python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c26_b.log 2>&1
cat gpurun_out/c26_cfg5.log; grep -v "^==" gpurun_out/c26_launches.csv | cut -d, -f5,13- | cut -c1-40,100-200 | head -14
