#!/bin/bash
# multi-GPU validation of round 2: bash tools/gpu_multi3.sh N tag [samples]
N=${1:-2}; TAG=${2:-r03}; SAMPLES=${3:-10000}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_multi_g${N}_$TAG.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# two-GPU tests of the command line (newref -gpus 2 over NCCL, -partfiles, testbatch -gpus 2)
timeout 900 python -m pytest tests/test_cli_gpu.py -q -x -k "two_gpus" 2>&1 | tail -3 | tee $OUT/pytest_cli_g${N}_$TAG.log
# parity of the sharded symmetric search under NCCL: small shape (whole table vs oracle), then the benchmark shape (sampled)
timeout 300 $TR --master-port 29513 tools/check_sharded_sym.py 1 61 100 2>&1 | tail -1 | tee $OUT/check_sharded_sym_g${N}_$TAG.log
timeout 600 $TR --master-port 29514 tools/check_sharded_sym.py 1 600 100 50000 2>&1 | tail -1 | tee -a $OUT/check_sharded_sym_g${N}_$TAG.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pc=d["config"].get("parity_check") or {}
    print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phases_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"],3) if d.get("e2e") else None, "parity", pc.get("identical"), pc.get("whole_table_equals_single_gpu_search"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
for SH in sym; do
F=$OUT/bench_newref_600x50kb_g${N}_${SH}_$TAG
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --shard $SH --no-test > $F.json 2> $F.err; summ $F.json; tail -2 $F.err
done
F=$OUT/bench_default_g${N}_$TAG
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $F.json 2> $F.err; summ $F.json; tail -2 $F.err
F=$OUT/bench_newref_2000x10kb_g${N}_$TAG
timeout 1200 $TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 1 --workload newref_2000x10kb > $F.json 2> $F.err; summ $F.json; tail -2 $F.err
timeout 1200 $TR --master-port 29516 tools/test_10k.py --samples $SAMPLES --oracle 0 > $OUT/test10k_g${N}_$TAG.json 2> $OUT/test10k_g${N}_$TAG.err; tail -c 1500 $OUT/test10k_g${N}_$TAG.json; tail -2 $OUT/test10k_g${N}_$TAG.err
