#!/bin/bash
# multi-GPU parity + bench lines: bash tools/gpu_multi2.sh N tag
N=${1:-2}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_multi_$TAG.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# parity of the sharded symmetric search under NCCL: small shape (whole table vs oracle), then the benchmark shape (sampled)
timeout 300 $TR --master-port 29513 tools/check_sharded_sym.py 1 61 100 2>&1 | tail -1 | tee $OUT/check_sharded_sym_g${N}_$TAG.log
timeout 600 $TR --master-port 29514 tools/check_sharded_sym.py 1 600 100 50000 2>&1 | tail -1 | tee -a $OUT/check_sharded_sym_g${N}_$TAG.log
for W in newref_600x50kb; do
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $W > $OUT/bench_${W}_g${N}_$TAG.json 2> $OUT/bench_${W}_g${N}_$TAG.err
tail -c 5000 $OUT/bench_${W}_g${N}_$TAG.json; tail -3 $OUT/bench_${W}_g${N}_$TAG.err
done
