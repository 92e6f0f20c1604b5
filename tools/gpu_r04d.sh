#!/bin/bash
TAG=${1:-r04d}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/k6_select_cmp.py newref_600x50kb newref_600x250kb newref_2000x10kb 2>&1 | tee $OUT/k6_select_$TAG.txt | cut -c1-330
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_test_gpu.py -q -k "min_effect" > $OUT/sanitizer_memcheck_segmin_$TAG.log 2>&1; tail -4 $OUT/sanitizer_memcheck_segmin_$TAG.log
