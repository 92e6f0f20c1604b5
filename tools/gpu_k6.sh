#!/bin/bash
TAG=${1:-r02p}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py tests/test_search_f16_gpu.py -q -x > $OUT/pytest_search_$TAG.log 2>&1; tail -4 $OUT/pytest_search_$TAG.log
timeout 120 python tools/profile_k5.py newref_600x50kb 0 2>&1 | tail -1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wc_finalize -s 2 -c 1 -o $OUT/k6_$TAG \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_k6_$TAG.log 2>&1
tail -1 $OUT/ncu_k6_$TAG.log | cut -c1-200
