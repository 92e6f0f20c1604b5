#!/bin/bash
TAG=${1:-r04g}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/k6_select_cmp.py newref_600x50kb newref_600x250kb newref_2000x10kb 2>&1 | tee $OUT/k6_select_$TAG.txt | cut -c1-330
