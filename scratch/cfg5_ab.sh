#!/bin/bash
# cfg5: parity tests of the stratified kernels, then the staged and the
# in-place form back to back on the same box (usage: cfg5_ab.sh [bench opts])
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_classify.py tests/test_distributed.py -m gpu -q -x -k "strat or cfg5 or merge" 2>&1 | tail -3
for o in "" "--opt strata_nopart=1" "$*"; do
  echo "== $o"
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_on_sample'))"
done
