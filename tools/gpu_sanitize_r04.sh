#!/bin/bash
# compute-sanitizer over the search (K4h, K5t all modes, K6 split form) and the test path: logs kept under profiles/
TAG=${1:-r04k}
OUT=gpurun_out; mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $SAN --tool memcheck --error-exitcode 1 python -m pytest tests/test_search_sym_gpu.py tests/test_search_gpu.py -q -x -k "ties or too_few or nan or medium_vs_c_oracle or symmetric_vs_c_oracle or partial_ranges or both_selects" > $OUT/sanitizer_memcheck_search_$TAG.log 2>&1; echo "memcheck search rc=$?"; tail -4 $OUT/sanitizer_memcheck_search_$TAG.log
timeout 900 $SAN --tool memcheck --error-exitcode 1 python -m pytest tests/test_test_gpu.py -q -x -k "golden or batch_vs_oracle or segmentation_vs_oracle or min_effect or non_finite" > $OUT/sanitizer_memcheck_test_$TAG.log 2>&1; echo "memcheck test rc=$?"; tail -3 $OUT/sanitizer_memcheck_test_$TAG.log
timeout 1200 $SAN --tool racecheck --error-exitcode 1 python -m pytest tests/test_search_sym_gpu.py -q -x -k "symmetric_vs_c_oracle or both_selects" > $OUT/sanitizer_racecheck_search_$TAG.log 2>&1; echo "racecheck search rc=$?"; tail -4 $OUT/sanitizer_racecheck_search_$TAG.log
