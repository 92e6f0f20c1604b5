#!/bin/bash
# round-2 profile set: K5t cycle counters, ncu --set full of the K5t passes and of the K6 kernels, DRAM/L2 bytes of one step
TAG=${1:-r03c}
OUT=gpurun_out; mkdir -p $OUT
W=newref_600x50kb
for opt in "k5_f16=2" "k5_f16=2 k5_group=1" "k5_f16=2 k5_group=4" "k5_f16=2 k5_sym=0"; do
  echo "== $opt"
  timeout 120 python tools/profile_k5.py $W 0 $opt 2>&1 | tail -1
done > $OUT/tc_prof_$TAG.txt 2>&1
cat $OUT/tc_prof_$TAG.txt | cut -c1-700
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wc_dist_topk_tc|wc_fin_" -s 6 -c 6 -o $OUT/k5t_k6_$TAG \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_k5t_k6_$TAG.log 2>&1
tail -2 $OUT/ncu_k5t_k6_$TAG.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 60 --csv --log-file $OUT/launches_default_$TAG.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_launches_$TAG.log 2>&1
tail -1 $OUT/ncu_launches_$TAG.log | cut -c1-200
