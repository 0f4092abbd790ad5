#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/c7_nccl_test.log 2>&1; tail -3 gpurun_out/c7_nccl_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c7_2gpu.json 2> gpurun_out/c7_2gpu.err
tail -c 1500 gpurun_out/c7_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c7_2gpu.json').read().strip().splitlines()[-1])
    print('cfg2', d['n_gpus'], d['value'], d['ms_per_step'], 'merged', d.get('parity_merged'), 'e2e', d['e2e']['value'])
    for k,v in d.get('extra',{}).items(): print(k, v['value'], round(v['ms_per_step'],3), round(v['roofline']['frac'],4), v.get('parity_on_sample'), v.get('parity_merged'))
except Exception as e: print('ERR', e)
PY
