#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/fin2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/r2_gpu_tests.log
timeout 900 python bench.py > $O/r2_bench_default_1gpu.json 2> $O/bench_default.err
timeout 900 python scratch/file_bench.py > $O/r2_file_bench.json 2> $O/file_bench.err
tail -2 $O/r2_gpu_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/fin2/r2_bench_default_1gpu.json').read().strip().splitlines()[-1])
print('cfg2', '%.3e'%d['value'], d['ms_per_step'], round(d['roofline']['frac'],4), 'e2e %.3e'%d['e2e']['value'])
for k,v in d['extra'].items(): print(k, '%.3e'%v['value'], round(v['ms_per_step'],3), round(v['roofline']['frac'],4), v['roofline']['kernel'], v.get('parity_on_sample'))
f=json.load(open('gpurun_out/fin2/r2_file_bench.json'))
print('file', f['device_reader']['records_per_s'], f['coords']['device_reader']['records_per_s'])
PY
