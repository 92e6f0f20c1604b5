#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_search_f16_gpu.py -q -x -k "tcgen05 or sharded or margin" > $OUT/pytest_tc_$TAG.log 2>&1; tail -5 $OUT/pytest_tc_$TAG.log
W=newref_600x50kb
for opt in "k5_f16=2" "k5_f16=2 k5_sym=0" "k5_f16=2 k5_group=1" "k5_f16=2 k5_group=4"; do
  echo "== $opt"
  timeout 120 python tools/profile_k5.py $W 0 $opt 2>&1 | tail -1
done > $OUT/tc_prof_$TAG.txt 2>&1
cat $OUT/tc_prof_$TAG.txt
