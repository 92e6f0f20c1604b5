#!/bin/bash
TAG=${1:-r03}
OUT=gpurun_out; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_peak tools/l2_peak.cu && (/tmp/l2_peak 48; /tmp/l2_peak 96) | tee $OUT/l2_peak_$TAG.json
