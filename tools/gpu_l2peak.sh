#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_peak tools/l2_peak.cu && (/tmp/l2_peak 48; /tmp/l2_peak 96; /tmp/l2_peak 24) | tee $OUT/l2_peak_r03.json
