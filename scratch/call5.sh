#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_classify.py -m gpu -x -q > gpurun_out/c5_tests.log 2>&1
tail -15 gpurun_out/c5_tests.log
rm -f gpurun_out/c5_bench.jsonl
for args in "" "--ranks phylum,genus,species" "--ranks phylum,genus,species --mode above" "--workload cfg4" "--mode above" "--ranks phylum,genus,species --mode uniq"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $args 2>&1 | tail -1 >> gpurun_out/c5_bench.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/c5_bench.jsonl'):
    try:
        d=json.loads(l); print(d['config'].get('ranks'), d['config'].get('mode'), d['config'].get('samples_per_gpu'), round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print('bad line', l[:300])
PY
