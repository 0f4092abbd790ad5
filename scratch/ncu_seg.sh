#!/bin/bash
# ncu capture of classify_seg_kernel on cfg2 (one launch), details to gpurun_out/
ncu --set full --clock-control none --import-source on -k regex:classify_seg_kernel -s 3 -c 1 \
    -o gpurun_out/seg_r1 -f python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_seg.log 2>&1
ncu -i gpurun_out/seg_r1.ncu-rep --page details > gpurun_out/seg_r1_details.txt 2>&1
