#!/bin/bash
cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/c20_b.log 2>&1
timeout 600 scratch/ncu_one.sh classify_strata_kernel cfg5_r2c --workload cfg5
grep -v "^==" gpurun_out/r2_launches_cfg5.csv | cut -d, -f5,12- | tail -30
