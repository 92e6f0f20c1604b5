#!/bin/bash
# two (or N) GPUs: whole GPU suite (the multi-GPU CLI tests run), sharded parity, bench lines for both partitions, configs[3] tool
N=${1:-2}; TAG=${2:-r02x}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_multi_$TAG.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_g${N}_$TAG.log 2>&1; tail -4 $OUT/pytest_gpu_g${N}_$TAG.log
timeout 600 $TR --master-port 29514 tools/check_sharded_sym.py 1 600 100 50000 2>&1 | tail -1 | tee $OUT/check_sharded_sym_g${N}_$TAG.log
for SH in rows sym; do
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --shard $SH --no-test > $OUT/bench_newref_600x50kb_g${N}_${SH}_$TAG.json 2> $OUT/bench_newref_600x50kb_g${N}_${SH}_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_newref_600x50kb_g${N}_${SH}_$TAG.json").read().strip().splitlines()[-1])
    print("$SH N=$N ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"], d["config"]["parity_check"].get("whole_table_equals_single_gpu_search"))
except Exception as e:
    print("$SH failed", e); print(open("$OUT/bench_newref_600x50kb_g${N}_${SH}_$TAG.err").read()[-1500:])
PY
done
timeout 600 $TR --master-port 29515 tools/test_10k.py --samples $((N * 1024)) --oracle 0 > $OUT/test10k_g${N}_$TAG.json 2> $OUT/test10k_g${N}_$TAG.err; tail -c 1500 $OUT/test10k_g${N}_$TAG.json; tail -2 $OUT/test10k_g${N}_$TAG.err
