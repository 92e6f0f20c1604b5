#!/bin/bash
timeout 600 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_f16_gpu.py -q -x 2>&1 | tail -2
bash tools/gpu_k6b.sh
