#!/bin/bash
# final multi-GPU lines of the round: bash tools/gpu_multi4.sh N tag
N=${1:-2}; TAG=${2:-r03}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pc=d["config"].get("parity_check") or {}
    t=d.get("test") or {}
    print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phases_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"],3) if d.get("e2e") else None, "parity", pc.get("identical"), pc.get("whole_table_equals_single_gpu_search"), "test", round(t.get("value",0)), round((t.get("e2e") or {}).get("value",0)))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_cli_gpu.py -q -x -k "two_gpus" 2>&1 | tail -2 | tee $OUT/pytest_cli_g${N}_$TAG.log; fi
F=$OUT/bench_default_g${N}_$TAG
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $F.json 2> $F.err; summ $F.json; tail -2 $F.err | cut -c1-300
F=$OUT/bench_newref_2000x10kb_g${N}_$TAG
timeout 1200 $TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 1 --workload newref_2000x10kb > $F.json 2> $F.err; summ $F.json; tail -2 $F.err | cut -c1-300
