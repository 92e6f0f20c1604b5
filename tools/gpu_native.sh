#!/bin/bash
# Verifies the PyTorch-free buffer backend: whole GPU suite (the command-line tests run on it by default), smoke, and the
# wall time of one `wisecondor.py test` invocation per backend on a 50 kb-shaped synthetic reference.
TAG=${1:-r01p}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; tail -15 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 600 python tools/cli_startup.py > $OUT/cli_startup_$TAG.json 2> $OUT/cli_startup_$TAG.err; cat $OUT/cli_startup_$TAG.json; tail -3 $OUT/cli_startup_$TAG.err
