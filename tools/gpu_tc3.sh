#!/bin/bash
TAG=${1:-r03d}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py tests/test_search_f16_gpu.py -q -x 2>&1 | tail -4 | tee $OUT/pytest_search_$TAG.log
W=newref_600x50kb
for opt in "k5_f16=2" "k5_f16=2 k5_sym=0"; do
  echo "== $opt"
  timeout 120 python tools/profile_k5.py $W 0 $opt 2>&1 | tail -1
done > $OUT/tc_prof_$TAG.txt 2>&1
cat $OUT/tc_prof_$TAG.txt | cut -c1-700
timeout 300 python tools/profile_k5.py newref_2000x10kb 0 k5_f16=2 2>&1 | tail -1 | cut -c1-700 | tee -a $OUT/tc_prof_$TAG.txt
