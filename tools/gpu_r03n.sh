#!/bin/bash
TAG=${1:-r03n}
OUT=gpurun_out; mkdir -p $OUT
# first-pass fraction of the symmetric search with the pivot pass in place
for f in 8 16 32 4; do echo "== k5_sym=$f"; timeout 120 python tools/profile_k5.py newref_600x50kb 0 k5_f16=2 k5_sym=$f 2>&1 | tail -1 | cut -c1-420; done | tee $OUT/tc_symfrac_$TAG.txt
bash tools/gpu_sanitize.sh $TAG
