#!/bin/bash
# Final pass of a round: full GPU suite, smoke, bench lines (both arms), launch lists and ncu captures of every hot kernel.
TAG=${1:-r01z}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -3 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; tail -c 1200 $OUT/bench_default_$TAG.json; tail -2 $OUT/bench_default_$TAG.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err; tail -c 800 $OUT/bench_reference_$TAG.json
timeout 300 python bench.py --steps 20 --warmup 3 --workload newref_600x250kb --no-cpu-baseline > $OUT/bench_newref_600x250kb_$TAG.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 120 --csv --log-file $OUT/launches_default_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
# symmetric search: a step launches K5 twice (threshold pass, symmetric pass) and K6 once: skip one step, capture the next
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wc_dist_topk|wc_finalize" -s 3 -c 3 -o $OUT/k5k6_50kb_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-test > $OUT/ncu_k5k6_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wc_zscore_kernel|wc_segment_kernel" -s 6 -c 6 -o $OUT/k8k9_$TAG -f \
    python tools/bench_testpath.py 50000 256 > $OUT/ncu_k8k9_$TAG.log 2>&1
ls -la $OUT | tail -12
