#!/bin/bash
# compute-sanitizer over the segment / sample / chunk / edge classify tests and
# the coverage tests (profiles/r2_sanitizer_*.log)
cd "$(dirname "$0")/.."
SEL='(seg or sample or chunk or edge or long or sink or randomised or none_without) and not 1e6 and not carry'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_classify.py tests/test_gpu_cover.py -m gpu -q -x -k "$SEL or cover or vectors or intervals or limits" \
  > gpurun_out/r2_san_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_classify.py tests/test_gpu_cover.py -m gpu -q -x -k "(seg or sample or edge or long) and not 1e6 and not carry or vectors or limits" \
  > gpurun_out/r2_san_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2_san_racecheck.log
tail -5 gpurun_out/r2_san_memcheck.log gpurun_out/r2_san_racecheck.log
