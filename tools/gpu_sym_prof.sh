#!/bin/bash
# Per-CTA cycle breakdown of the symmetric K5 (second pass) under a few settings.
OUT=gpurun_out; mkdir -p $OUT
W=${1:-newref_600x50kb}
for opt in "k5_sym=0" "k5_sym=8" "k5_sym=8 k5_group=1" "k5_sym=8 k5_group=4" "k5_sym=4" "k5_sym=16 k5_group=4" "k5_sym=8 k5_stages=5"; do
  echo "== $opt"
  timeout 120 python tools/profile_k5.py $W 0 $opt 2>&1 | tail -1
done > $OUT/sym_prof_r01q.txt 2>&1
cat $OUT/sym_prof_r01q.txt
