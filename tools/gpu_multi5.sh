#!/bin/bash
# the default bench line at N GPUs (final code of the round): bash tools/gpu_multi5.sh N tag
N=${1:-8}; TAG=${2:-r03}
OUT=gpurun_out; mkdir -p $OUT
F=$OUT/bench_default_g${N}_$TAG
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $F.json 2> $F.err
python - "$F.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
pc=d["config"].get("parity_check") or {}; t=d.get("test") or {}
print("ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phases_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"],3), "parity", pc.get("identical"), pc.get("whole_table_equals_single_gpu_search"), "test", round(t.get("value",0)), round((t.get("e2e") or {}).get("value",0)))
PY
tail -2 $F.err | cut -c1-200
