#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_ordinal.py tests/test_gpu_classify.py -m gpu -q -x 2>&1 | tail -2
for o in "" "--opt cnt_nowin=1" ""; do echo "== $o"; timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_on_sample'))"; done
scratch/launches.sh cfg3 | tail -7
