#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_classify.py -m gpu -q -x -k "strat or cfg5" 2>&1 | tail -12 > gpurun_out/c23_tests.log
for o in "" "--opt strata_nopart=1"; do
  echo "== $o" >> gpurun_out/c23_cfg5.log
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity'), d.get('checksum'))" >> gpurun_out/c23_cfg5.log 2>&1
done
tail -6 gpurun_out/c23_tests.log; cat gpurun_out/c23_cfg5.log
