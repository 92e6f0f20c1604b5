#!/bin/bash
# Round-end pass after the symmetric search became the default: GPU suite, smoke, default bench, launch list, and the DRAM
# traffic + duration of the two K5 launches (metrics-only ncu pass: cheap).
TAG=${1:-r01s}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -6 $OUT/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 400 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; head -c 1800 $OUT/bench_default_$TAG.json; echo; tail -2 $OUT/bench_default_$TAG.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
    -k regex:"wc_dist_topk|wc_finalize|wc_prepare" -s 3 -c 4 --csv --log-file $OUT/k5_dram_sym_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-test > $OUT/ncu_dram_$TAG.log 2>&1
grep -v "^==" $OUT/k5_dram_sym_$TAG.csv | cut -d, -f5,12-15 | tail -20
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv --log-file $OUT/launches_default_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-test > $OUT/ncu_launch_$TAG.log 2>&1
tail -3 $OUT/launches_default_$TAG.csv | cut -c1-200
