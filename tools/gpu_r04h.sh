#!/bin/bash
# wc_newref_topk_host: K6 in 1 / 2 / 4 / 8 row ranges (end-to-end time of the default workload), then the host-call tests
TAG=${1:-r04h}
OUT=gpurun_out; mkdir -p $OUT
for P in 2 4 8 1 4; do
WC_K6_PARTS=$P timeout 300 python bench.py --steps 10 --warmup 3 --no-test --no-cpu-baseline 2>$OUT/bench_parts_$TAG.err | tee $OUT/bench_parts${P}_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parts $P ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['config']['parity_check']['identical'])"
done
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_f16_gpu.py tests/test_cli_gpu.py tests/test_native_backend_gpu.py -m gpu -q 2>&1 | tail -3
