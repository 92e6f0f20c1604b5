#!/bin/bash
# whole GPU suite + smoke + default bench line: bash tools/gpu_suite.sh tag
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -8 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -3 $OUT/smoke_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; tail -c 1500 $OUT/bench_default_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_default_$TAG.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
t = d.get("test") or {}
print("test", t.get("value"), t.get("phases_ms"), (t.get("e2e") or {}).get("value"))
PY
