#!/bin/bash
cd "$(dirname "$0")/.."
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_${N}gpu.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_bench_{N}gpu.json').read().strip().splitlines()[-1])
    print('cfg2', d['n_gpus'], '%.3e'%d['value'], d['ms_per_step'], 'merged', d.get('parity_merged'), 'e2e %.3e'%d['e2e']['value'], 'soa %.3e'%d['e2e']['int32_soa']['value'])
    for k,v in d.get('extra',{}).items(): print(k, '%.3e'%v['value'], round(v['ms_per_step'],3), round(v['roofline']['frac'],4), v.get('parity_on_sample'), v.get('parity_merged'))
except Exception as e: print('ERR', e)
PY
