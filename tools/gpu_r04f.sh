#!/bin/bash
# two GPUs with the round's final code: the two-GPU CLI tests, sharded parity, bench lines for both partitions
N=${1:-2}; TAG=${2:-r04f}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_search_shard_gpu.py -m gpu -q > $OUT/pytest_cli_g${N}_$TAG.log 2>&1; tail -3 $OUT/pytest_cli_g${N}_$TAG.log
timeout 300 $TR --master-port 29514 tools/check_sharded_sym.py 1 600 100 50000 2>&1 | tail -1 | tee $OUT/check_sharded_sym_g${N}_$TAG.log | cut -c1-400
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_default_g${N}_$TAG.json 2> $OUT/bench_default_g${N}_$TAG.err
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --shard sym --no-test --no-cpu-baseline > $OUT/bench_newref_600x50kb_g${N}_sym_$TAG.json 2> $OUT/bench_newref_600x50kb_g${N}_sym_$TAG.err
for F in bench_default_g${N}_$TAG bench_newref_600x50kb_g${N}_sym_$TAG; do
python - <<PY
import json
try:
    d = json.loads(open("$OUT/$F.json").read().strip().splitlines()[-1])
    print("$F ms/step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["phases_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"], d["config"]["parity_check"].get("whole_table_equals_single_gpu_search"), "test", round(d.get("test", {}).get("value", 0)))
except Exception as e:
    print("$F failed", e); print(open("$OUT/$F.err").read()[-1500:])
PY
done
