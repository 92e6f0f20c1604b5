#!/bin/bash
# multi-GPU bench lines: bash tools/gpu_multi.sh N tag
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_multi_$TAG.txt
# parity of both sharded searches under NCCL first
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/check_sharded_sym.py 2>&1 | tail -2
for W in newref_600x50kb newref_600x250kb; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload $W > $OUT/bench_${W}_g${N}_$TAG.json 2> $OUT/bench_${W}_g${N}_$TAG.err
tail -c 3500 $OUT/bench_${W}_g${N}_$TAG.json; tail -3 $OUT/bench_${W}_g${N}_$TAG.err
done
