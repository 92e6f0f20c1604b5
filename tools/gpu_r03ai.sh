#!/bin/bash
TAG=${1:-r03ai}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_test_gpu.py tests/test_cli_gpu.py -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_nocpu_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['test']; print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'test', round(t['value']), t['phases_ms'], 'e2e', round(t['e2e']['value']))"
