#!/bin/bash
TAG=${1:-r03s}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py -q -x 2>&1 | tail -3
timeout 600 python tools/k6_sweep.py newref_600x50kb 4,2,100 2>&1 | cut -c1-330 | tee $OUT/k6_sweep_50kb_$TAG.txt
timeout 900 python tools/k6_sweep.py newref_2000x10kb 4,2,100 2>&1 | cut -c1-330 | tee $OUT/k6_sweep_10kb_$TAG.txt
bash tools/gpu_k6_launches.sh $TAG 2>&1 | tail -14
