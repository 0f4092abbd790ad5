#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c17_tests.log
timeout 400 python scratch/file_prof.py > gpurun_out/c17_prof.log 2>&1
timeout 600 python scratch/file_bench.py > gpurun_out/c17_file_bench.json 2> gpurun_out/c17_file_bench.err
tail -4 gpurun_out/c17_tests.log; cat gpurun_out/c17_file_bench.json; tail -3 gpurun_out/c17_file_bench.err
