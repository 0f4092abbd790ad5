#!/bin/bash
# reader options on the device + where the file path's host time goes
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parse.py tests/test_dropin_reference.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c14_tests.log
timeout 300 python scratch/file_prof.py > gpurun_out/c14_prof.log 2>&1
tail -5 gpurun_out/c14_tests.log
