#!/bin/bash
cd "$(dirname "$0")/.."
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/c6_default.json 2> gpurun_out/c6_default.err
tail -c 600 gpurun_out/c6_default.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c6_default.json').read().strip().splitlines()[-1])
    print('cfg2', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['parity_on_sample'])
    for k,v in d.get('extra',{}).items(): print(k, round(v['ms_per_step'],3), round(v['roofline']['frac'],4), v['roofline']['kernel'], v.get('parity_on_sample'))
except Exception as e: print('ERR', e)
PY
