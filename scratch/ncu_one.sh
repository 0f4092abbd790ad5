#!/bin/bash
# usage: ncu_one.sh <kernel-regex> <out-name> <bench args...>: one `ncu --set full` capture
cd "$(dirname "$0")/.."
k=$1; o=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/$o \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/$o.log 2>&1
tail -2 gpurun_out/$o.log
