#!/bin/bash
TAG=${1:-r03o}
OUT=gpurun_out; mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py -q -x 2>&1 | tail -2
timeout 1200 $SAN --tool racecheck --error-exitcode 1 python -m pytest tests/test_search_sym_gpu.py -q -x -k "symmetric_vs_c_oracle" > $OUT/sanitizer_racecheck_search_$TAG.log 2>&1; echo "racecheck search rc=$?"; grep -v "^=========     and" $OUT/sanitizer_racecheck_search_$TAG.log | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-test --no-cpu-baseline 2>/dev/null | tee $OUT/bench_quick_newref_600x50kb_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],3), d['phases_ms'], 'roofline', d['roofline']['kernel'][:24], round(d['roofline']['frac'],3), d['roofline'].get('traffic'), '| 2nd', d['roofline_second_kernel']['kernel'][:24], round(d['roofline_second_kernel']['frac'],3), d['roofline_second_kernel'].get('traffic'), d['roofline_second_kernel'].get('l2'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wc_dist_topk_tc|wc_fin_" -s 6 -c 6 -o $OUT/k5t_k6_$TAG \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_k5t_k6_$TAG.log 2>&1
tail -1 $OUT/ncu_k5t_k6_$TAG.log | cut -c1-100
