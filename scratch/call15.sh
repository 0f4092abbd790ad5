#!/bin/bash
# reader options + block reader on the device; file path timing
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parse.py tests/test_dropin_reference.py -m gpu -q 2>&1 | tail -40 > gpurun_out/c15_tests.log
timeout 300 python scratch/file_prof.py > gpurun_out/c15_prof.log 2>&1
timeout 300 python scratch/file_bench.py > gpurun_out/c15_file_bench.json 2> gpurun_out/c15_file_bench.err
tail -5 gpurun_out/c15_tests.log; cat gpurun_out/c15_file_bench.json
