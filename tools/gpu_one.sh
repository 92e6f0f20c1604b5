#!/bin/bash
TAG=${1:-r02w}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_f16_gpu.py tests/test_test_gpu.py -q -x > $OUT/pytest_part_$TAG.log 2>&1; tail -3 $OUT/pytest_part_$TAG.log
timeout 600 python tools/test_10k.py --samples 1024 --oracle 1 > $OUT/test10k_g1_$TAG.json 2> $OUT/test10k_g1_$TAG.err; tail -c 1800 $OUT/test10k_g1_$TAG.json; tail -3 $OUT/test10k_g1_$TAG.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-test > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; tail -c 600 $OUT/bench_default_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_default_$TAG.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), d["phases_ms"], "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["config"]["parity_check"]["identical"])
print("roofline", d["roofline"]["kernel"][:40], round(d["roofline"]["frac"],3), "| second", d["roofline_second_kernel"]["kernel"][:40], round(d["roofline_second_kernel"]["frac"],3))
PY
