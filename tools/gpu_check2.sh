#!/bin/bash
# smoke + default bench + ncu captures of K5 (50 kb), K8 and K9
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -3 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; tail -c 4000 $OUT/bench_default_$TAG.json; tail -3 $OUT/bench_default_$TAG.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wc_dist_topk -s 1 -c 1 -o $OUT/k5_50kb_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-test > $OUT/ncu_k5_50kb_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wc_zscore_kernel -s 5 -c 2 -o $OUT/k8_$TAG -f \
    python tools/bench_testpath.py 50000 256 > $OUT/ncu_k8_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wc_segment_kernel -s 1 -c 1 -o $OUT/k9_$TAG -f \
    python tools/bench_testpath.py 50000 256 > $OUT/ncu_k9_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file $OUT/launches_test_$TAG.csv \
    python tools/bench_testpath.py 50000 256 > $OUT/ncu_launch_test_$TAG.log 2>&1
ls -la $OUT | tail -20
