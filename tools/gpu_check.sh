#!/bin/bash
# One GPU-box pass: parity tests, bench lines, ncu launch list and one full capture of the distance kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_50kb_$TAG.json 2> $OUT/bench_50kb_$TAG.err; tail -c 3000 $OUT/bench_50kb_$TAG.json
timeout 300 python bench.py --steps 20 --warmup 3 --workload newref_600x250kb --no-cpu-baseline > $OUT/bench_250kb_$TAG.json 2> $OUT/bench_250kb_$TAG.err; tail -c 1500 $OUT/bench_250kb_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wc_dist_topk -s 1 -c 1 -o $OUT/k5_$TAG -f \
    python bench.py --steps 1 --warmup 3 --workload newref_600x250kb --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
