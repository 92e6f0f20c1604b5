#!/bin/bash
TAG=${1:-r03ab}
OUT=gpurun_out; mkdir -p $OUT
for g in 0 1 2 4 8; do echo "== k5_group=$g"; timeout 300 python tools/profile_k5.py newref_2000x10kb 0 k5_f16=2 k5_group=$g 2>&1 | tail -1 | cut -c1-330; done | tee $OUT/tc_group_10kb_$TAG.txt
for g in 1 2 4 8; do echo "== k5_group=$g"; timeout 120 python tools/profile_k5.py newref_600x50kb 0 k5_f16=2 k5_group=$g 2>&1 | tail -1 | cut -c1-330; done | tee $OUT/tc_group_50kb_$TAG.txt
