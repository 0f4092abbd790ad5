#!/bin/bash
# compute-sanitizer over the GPU parity tests of every kernel of the round
# (profiles/r2_sanitizer_*.log); the 1e6-record and carry cases are left out for time
cd "$(dirname "$0")/.."
SEL='not 1e6 and not carry and not at_1e6'
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_classify.py tests/test_gpu_cover.py tests/test_gpu_ordinal.py -m gpu -q -x -k "$SEL" \
  > gpurun_out/r2_san_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_san_memcheck.log
timeout 1100 compute-sanitizer --tool racecheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_classify.py tests/test_gpu_cover.py tests/test_gpu_ordinal.py -m gpu -q -x \
  -k "(modes_small or long_queries or samples_and_chunks or strata or edge_inputs or stratified_one or packed or contiguous_samples or fused or vectors or limits or which_kernel) and not 4096" \
  > gpurun_out/r2_san_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2_san_racecheck.log
tail -n 4 gpurun_out/r2_san_memcheck.log gpurun_out/r2_san_racecheck.log
