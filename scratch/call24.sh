#!/bin/bash
cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"strata|fill_slots" -c 80 --csv --log-file gpurun_out/c24_launches.csv python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c24_b.log 2>&1
grep -v "^==" gpurun_out/c24_launches.csv | cut -d, -f5,13- | head -150 | awk 'NR%1==0' | cut -c1-150
