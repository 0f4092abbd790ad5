#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/c36_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c36_cfg5_2gpu.json 2> gpurun_out/c36_err.log
tail -5 gpurun_out/c36_tests.log; python -c "
import json; d=json.loads(open('gpurun_out/c36_cfg5_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('parity_merged'), d['roofline']['kernel_ms'])"; grep -v Warning gpurun_out/c36_err.log | tail -5
