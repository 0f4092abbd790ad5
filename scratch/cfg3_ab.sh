#!/bin/bash
# cfg3: parity tests of the coordinate path, then the bench twice
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_ordinal.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_on_sample'))"; done
