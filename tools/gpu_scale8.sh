#!/bin/bash
# 8-GPU validation: default workload at N=8, configs[4] (2000 x 10 kb) at N=8, parity at that size on one GPU
TAG=${1:-r01}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 "${@:2}"; }
run 600 --steps 10 --warmup 3 > $OUT/bench_newref_600x50kb_g8_$TAG.json 2> $OUT/bench_newref_600x50kb_g8_$TAG.err; tail -c 3000 $OUT/bench_newref_600x50kb_g8_$TAG.json; tail -2 $OUT/bench_newref_600x50kb_g8_$TAG.err
run 900 --steps 3 --warmup 1 --workload newref_2000x10kb > $OUT/bench_newref_2000x10kb_g8_$TAG.json 2> $OUT/bench_newref_2000x10kb_g8_$TAG.err; tail -c 2500 $OUT/bench_newref_2000x10kb_g8_$TAG.json; tail -2 $OUT/bench_newref_2000x10kb_g8_$TAG.err
timeout 900 python tools/check_large.py 10000 2000 32 > $OUT/check_large_$TAG.json 2> $OUT/check_large_$TAG.err; cat $OUT/check_large_$TAG.json; tail -3 $OUT/check_large_$TAG.err
