#!/bin/bash
# Final pass of round 2, one GPU: the GPU suite, smoke, the default bench line
TAG=${1:-r04i}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 400 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; tail -2 $OUT/bench_default_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_default_$TAG.json").read().strip().splitlines()[-1])
t = d["test"]
print("ms/step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["phases_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"], "roofline", round(d["roofline"]["frac"], 3), round(d["roofline_second_kernel"]["frac"], 3), "test", round(t["value"]), round(t["e2e"]["value"]), "cpu", d["cpu_baseline"]["value"], d["clocks"])
PY
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_$TAG.json 2>/dev/null; cut -c1-300 $OUT/bench_reference_$TAG.json
