#!/bin/bash
# Round-2 end pass, one GPU: smoke, default bench line, launch list of the same command, ncu --set full of the new select
# kernel and the DRAM / L2 traffic of the K6 kernels.
TAG=${1:-r04e}
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 400 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; head -c 600 $OUT/bench_default_$TAG.json; echo; tail -2 $OUT/bench_default_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_default_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-test > $OUT/ncu_launch_$TAG.log 2>&1
tail -3 $OUT/launches_default_$TAG.csv | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
    -k regex:"wc_fin_|wc_dist_topk_tc" -s 12 -c 12 --csv --log-file $OUT/k6_traffic_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_traffic_$TAG.log 2>&1
grep -v "^==" $OUT/k6_traffic_$TAG.csv | cut -d, -f5,12-15 | tail -24
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wc_fin_select_hist -s 3 -c 1 -o $OUT/k6a_hist_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_k6a_$TAG.log 2>&1
tail -2 $OUT/ncu_k6a_$TAG.log | cut -c1-200
ncu -i $OUT/k6a_hist_$TAG.ncu-rep --page raw --csv > $OUT/k6a_hist_ncu_full_$TAG.csv 2>/dev/null; wc -c $OUT/k6a_hist_ncu_full_$TAG.csv
