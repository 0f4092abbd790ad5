#!/bin/bash
cd /root/repo
timeout 600 scratch/ncu_one.sh strata_apply_kernel cfg5_apply --workload cfg5
