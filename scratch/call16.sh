#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parse.py -m gpu -q 2>&1 | tail -30 > gpurun_out/c16_tests.log
timeout 400 python scratch/file_prof.py > gpurun_out/c16_prof.log 2>&1
timeout 600 python scratch/file_bench.py > gpurun_out/c16_file_bench.json 2> gpurun_out/c16_file_bench.err
tail -4 gpurun_out/c16_tests.log; cat gpurun_out/c16_file_bench.json; tail -3 gpurun_out/c16_file_bench.err
