#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_classify.py -m gpu -q -x -k "strat or cfg5" 2>&1 | tail -8 > gpurun_out/c19_tests.log
for o in "" "--opt strata_nowin=1" "--opt strata_wt=512" "--opt strata_wt=512 --opt strata_nowin=1" "--opt strata_nt=512" "--opt strata_nt=1024" "--opt strata_nt=640"; do
  echo "== $o" >> gpurun_out/c19_cfg5.log
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline'])" >> gpurun_out/c19_cfg5.log 2>&1
done
tail -3 gpurun_out/c19_tests.log; cat gpurun_out/c19_cfg5.log
