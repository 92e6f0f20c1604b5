#!/bin/bash
# 2000 x 10 kb on ONE GPU (dress rehearsal of the 8-GPU line): bench with the sampled-row parity check
TAG=${1:-r02y}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python bench.py --steps 3 --warmup 1 --workload newref_2000x10kb > $OUT/bench_newref_2000x10kb_g1_$TAG.json 2> $OUT/bench_newref_2000x10kb_g1_$TAG.err
tail -c 800 $OUT/bench_newref_2000x10kb_g1_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_newref_2000x10kb_g1_$TAG.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "parity", d["config"]["parity_check"])
print("roofline", d["roofline"]["kernel"][:40], round(d["roofline"]["frac"],3), "| second", d["roofline_second_kernel"]["kernel"][:40], round(d["roofline_second_kernel"]["frac"],3))
PY
