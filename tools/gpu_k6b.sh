#!/bin/bash
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
for W in newref_600x50kb newref_2000x10kb; do
timeout 600 python bench.py --steps 3 --warmup 1 --workload $W --no-test --no-cpu-baseline --no-parity-check 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W', round(d['ms_per_step'],3), d['phases_ms'])"
done
