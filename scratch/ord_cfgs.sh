#!/bin/bash
run() { echo -n "$1: "; env $1 python bench.py --workload cfg3 --no-cpu --no-e2e --steps 5 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), d["roofline"]["kernel_ms"])'; }
run "X=1"
run "WK_ORD_NOWIN=1"
