#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parse.py tests/test_dropin_reference.py tests/test_gpu_ordinal.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c18_tests.log
timeout 400 python scratch/file_prof.py > gpurun_out/c18_prof.log 2>&1
timeout 900 python scratch/file_bench.py > gpurun_out/c18_file_bench.json 2> gpurun_out/c18_file_bench.err
tail -4 gpurun_out/c18_tests.log; cat gpurun_out/c18_file_bench.json; tail -3 gpurun_out/c18_file_bench.err
