#!/bin/bash
# K5t experiments with tight time limits (a deadlock must not burn the budget)
TAG=${1:-r03ah}
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m pytest tests/test_search_sym_gpu.py -q -x 2>&1 | tail -3
for opt in "k5_f16=2" "k5_f16=2 k5_sym=0"; do
  echo "== $opt"
  timeout 90 python tools/profile_k5.py newref_600x50kb 0 $opt 2>&1 | tail -1
done > $OUT/tc_prof_$TAG.txt 2>&1
cat $OUT/tc_prof_$TAG.txt | cut -c1-420
