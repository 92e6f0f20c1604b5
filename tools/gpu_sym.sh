#!/bin/bash
# Symmetric search: parity tests, then K5/K6 times of the plain and the symmetric search on the default workload.
TAG=${1:-r01q}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_search_sym_gpu.py -q -x > $OUT/pytest_sym_$TAG.log 2>&1; tail -15 $OUT/pytest_sym_$TAG.log
for f in 8 16 32; do
  WC_K5_SYM=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-test > $OUT/bench_sym${f}_$TAG.json 2> $OUT/bench_sym${f}_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_sym${f}_$TAG.json").read().strip().splitlines()[-1])
    print("sym=$f", "ms/step", round(d["ms_per_step"], 2), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 2))
except Exception as e:
    print("sym=$f failed", e); print(open("$OUT/bench_sym${f}_$TAG.err").read()[-1500:])
PY
done
for f in 0 8; do
  WC_K5_SYM=$f timeout 200 python bench.py --steps 20 --warmup 3 --workload newref_600x250kb --no-cpu-baseline --no-test > $OUT/bench_250kb_sym${f}_$TAG.json 2> /dev/null
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_250kb_sym${f}_$TAG.json").read().strip().splitlines()[-1])
    print("250kb sym=$f", "ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("250kb sym=$f failed", e)
PY
done
