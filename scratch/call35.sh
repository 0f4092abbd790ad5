#!/bin/bash
cd /root/repo
timeout 600 python scratch/merge_bench.py > gpurun_out/c35_merge.log 2>&1
tail -40 gpurun_out/c35_merge.log
