#!/bin/bash
# cfg5 at N GPUs (default 2): the job with the reduce-scatter merge
cd "$(dirname "$0")/.."
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('parity_merged'), d['roofline']['kernel_ms'])"
