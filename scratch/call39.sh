#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ordinal.py tests/test_distributed.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/c39_tests.log
for o in "" "--opt strata_nopart=1" ""; do
  echo "== $o" >> gpurun_out/c39_cfg5.log
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_on_sample'))" >> gpurun_out/c39_cfg5.log 2>&1
done
timeout 300 python scratch/merge_bench.py 2>&1 | tail -11 >> gpurun_out/c39_cfg5.log
cat gpurun_out/c39_tests.log gpurun_out/c39_cfg5.log
