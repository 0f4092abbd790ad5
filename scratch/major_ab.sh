#!/bin/bash
# --major through classify_multi_kernel: parity, then cfg4 --major 80 with and
# without it
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_classify.py tests/test_golden.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -4
for o in ""; do
  echo "== $o"
  timeout 300 python bench.py --workload cfg4 --ranks phylum,genus,species --mode major --samples 8 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d.get('parity_on_sample'))"
done
timeout 300 python bench.py --ranks genus --mode major --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('genus major', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d.get('parity_on_sample'))"
