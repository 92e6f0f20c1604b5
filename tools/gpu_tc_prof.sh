#!/bin/bash
# Where K5t's time goes: per-CTA cycle counters under a few settings, then one ncu --set full capture of both launches.
TAG=${1:-r02d}
OUT=gpurun_out; mkdir -p $OUT
W=newref_600x50kb
for opt in "k5_f16=2" "k5_f16=2 k5_sym=0" "k5_f16=2 k5_group=1" "k5_f16=2 k5_group=4" "k5_f16=1"; do
  echo "== $opt"
  timeout 120 python tools/profile_k5.py $W 0 $opt 2>&1 | tail -1
done > $OUT/tc_prof_$TAG.txt 2>&1
cat $OUT/tc_prof_$TAG.txt
WC_K5_F16=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wc_dist_topk_tc -s 2 -c 2 -o $OUT/k5t_$TAG \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_k5t_$TAG.log 2>&1
tail -3 $OUT/ncu_k5t_$TAG.log
