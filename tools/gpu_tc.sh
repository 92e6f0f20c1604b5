#!/bin/bash
# The tcgen05 / TMEM filter (k5_f16 = 2): error-vs-margin test first (it also proves the descriptors), parity, bench lines.
TAG=${1:-r02c}
OUT=gpurun_out; mkdir -p $OUT
timeout 180 python -m pytest tests/test_search_f16_gpu.py -q -x -s -k "margin" > $OUT/pytest_tc_margin_$TAG.log 2>&1; tail -12 $OUT/pytest_tc_margin_$TAG.log
timeout 600 python -m pytest tests/test_search_f16_gpu.py -q -x -k "tcgen05 or sharded" > $OUT/pytest_tc_$TAG.log 2>&1; tail -12 $OUT/pytest_tc_$TAG.log
for W in newref_600x50kb newref_600x250kb; do
  WC_K5_F16=2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $W --no-cpu-baseline --no-test > $OUT/bench_${W}_tc_$TAG.json 2> $OUT/bench_${W}_tc_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${W}_tc_$TAG.json").read().strip().splitlines()[-1])
    print("$W tc", "ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"])
except Exception as e:
    print("$W tc failed", e); print(open("$OUT/bench_${W}_tc_$TAG.err").read()[-1500:])
PY
done
