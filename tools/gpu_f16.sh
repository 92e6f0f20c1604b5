#!/bin/bash
# The experimental fp16 tensor-core filter (k5_f16): parity tests, then bench lines with and without it.
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
WC_TEST_F16=1 timeout 900 python -m pytest tests/test_search_f16_gpu.py -q -x > $OUT/pytest_f16_$TAG.log 2>&1; tail -15 $OUT/pytest_f16_$TAG.log
for f in 0 1; do
  for W in newref_600x50kb newref_600x250kb; do
    WC_K5_F16=$f timeout 300 python bench.py --steps 5 --warmup 3 --workload $W --no-cpu-baseline --no-test > $OUT/bench_${W}_f16_${f}_$TAG.json 2> $OUT/bench_${W}_f16_${f}_$TAG.err
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${W}_f16_${f}_$TAG.json").read().strip().splitlines()[-1])
    print("$W f16=$f", "ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("$W f16=$f failed", e); print(open("$OUT/bench_${W}_f16_${f}_$TAG.err").read()[-1500:])
PY
  done
done
