#!/bin/bash
TAG=${1:-r03v}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/k6_sweep.py newref_600x50kb 4,2,100,0 4,2,86,1 4,1,86,1 4,4,86,1 4,2,76,1 4,2,60,1 8,1,40,1 8,2,40,1 2>&1 | cut -c1-260 | tee $OUT/k6_sweep_50kb_$TAG.txt
timeout 600 python tools/k6_sweep.py newref_2000x10kb 4,2,100,0 4,2,86,1 2>&1 | cut -c1-260 | tee $OUT/k6_sweep_10kb_$TAG.txt
