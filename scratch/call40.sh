#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_classify.py tests/test_dropin_reference.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/c40_tests.log
timeout 600 python bench.py --no-extra --no-cpu > gpurun_out/c40_bench.json 2> gpurun_out/c40_err.log
cat gpurun_out/c40_tests.log; python -c "
import json; d=json.loads(open('gpurun_out/c40_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])[:900])"; tail -3 gpurun_out/c40_err.log
