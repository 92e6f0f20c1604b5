#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out; mkdir -p $OUT
WC_K5_F16=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wc_dist_topk_tc -s 2 -c 2 -o $OUT/k5t_$TAG \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_k5t_$TAG.log 2>&1
tail -2 $OUT/ncu_k5t_$TAG.log | cut -c1-300
