#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_classify.py -m gpu -q -x -k "strat or cfg5" 2>&1 | tail -3 > gpurun_out/c27_tests.log
for o in "" "--opt strata_dbg=4" "--opt strata_dbg=8" "--opt strata_bpp=1184" "--opt strata_bpp=1184 --opt strata_dbg=8"; do
  echo "== $o" >> gpurun_out/c27_cfg5.log
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" >> gpurun_out/c27_cfg5.log 2>&1
done
cat gpurun_out/c27_tests.log gpurun_out/c27_cfg5.log
