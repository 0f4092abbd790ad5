#!/bin/bash
# compute-sanitizer over what came after r2_sanitize.sh: the staged strata
# kernels (classify_strata_kernel<..,256>, strata_apply_kernel) and the device
# reader with its options (wk_parse.cuh) -> profiles/r2_sanitizer_late_*.log
cd "$(dirname "$0")/.."
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_parse.py tests/test_gpu_classify.py -m gpu -q -x \
  -k "stratified_one or strata or test_gpu_parse" > gpurun_out/r2_san_late_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_san_late_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_parse.py tests/test_gpu_classify.py -m gpu -q -x \
  -k "stratified_one or trim_sub or coordinates or odd_lines or mates" > gpurun_out/r2_san_late_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2_san_late_racecheck.log
tail -n 5 gpurun_out/r2_san_late_memcheck.log gpurun_out/r2_san_late_racecheck.log
