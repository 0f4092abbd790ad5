#!/bin/bash
cd "$(dirname "$0")/.."
bash scratch/ncu_one.sh "classify_kernel" cfg5_r2a --workload cfg5
bash scratch/ncu_one.sh "ordinal_match" cfg3_r2a --workload cfg3
ls -la gpurun_out
