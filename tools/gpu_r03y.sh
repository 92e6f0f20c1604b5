#!/bin/bash
TAG=${1:-r03y}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_test_gpu.py -q -x 2>&1 | tail -2
timeout 400 python tools/cli_profile.py 2>&1 | tail -48 | cut -c1-150 | tee $OUT/cli_profile_$TAG.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_nocpu_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['test']; print('ms/step', round(d['ms_per_step'],3), 'test', round(t['value']), t['phases_ms'], 'e2e', round(t['e2e']['value']), 'K9 frac', round(t['roofline']['K9']['frac'],3))"
