#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_classify.py -m gpu -x -q > gpurun_out/c2_tests.log 2>&1
tail -5 gpurun_out/c2_tests.log
for args in "" "--mode uniq" "--ranks phylum,genus,species" "--ranks none"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $args 2>&1 | tail -1 >> gpurun_out/c2_bench.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/c2_bench.jsonl'):
    try:
        d=json.loads(l); print(d['config'].get('ranks'), d['config'].get('mode'), round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print('bad line', l[:200])
PY
