#!/bin/bash
TAG=${1:-r03ac}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py tests/test_search_f16_gpu.py tests/test_cli_gpu.py -q -x 2>&1 | tail -3
for W in newref_600x50kb newref_600x250kb; do
timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-test --no-cpu-baseline 2>$OUT/bench_quick_$TAG.err | tee $OUT/bench_quick_${W}_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W', 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['config']['parity_check']['identical'])"
done
