#!/bin/bash
# N-GPU weak-scaling line of the default bench; usage: multi.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu 2>&1 | tail -1 | tee gpurun_out/r1_bench_cfg2_${N}gpu.json
