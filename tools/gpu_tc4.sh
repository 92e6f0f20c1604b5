#!/bin/bash
TAG=${1:-r03f}
OUT=gpurun_out; mkdir -p $OUT
W=newref_600x50kb
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py -q -x 2>&1 | tail -3
for opt in "k5_f16=2" "k5_f16=2 k5_sym=0"; do
  echo "== $opt"
  timeout 120 python tools/profile_k5.py $W 0 $opt 2>&1 | tail -1
done > $OUT/tc_prof_$TAG.txt 2>&1
cat $OUT/tc_prof_$TAG.txt | cut -c1-900
