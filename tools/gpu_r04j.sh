#!/bin/bash
# N GPUs, final code of round 2: the default bench line as the driver launches it
N=${1:-4}; TAG=${2:-r04j}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_default_g${N}_$TAG.json 2> $OUT/bench_default_g${N}_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_default_g${N}_$TAG.json").read().strip().splitlines()[-1])
print("N=$N ms/step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["phases_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"], d["config"]["parity_check"].get("whole_table_equals_single_gpu_search"), "test", round(d["test"]["value"]), round(d["test"]["e2e"]["value"]))
PY
