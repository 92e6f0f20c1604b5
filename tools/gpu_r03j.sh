#!/bin/bash
TAG=${1:-r03j}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_search_sym_gpu.py tests/test_search_shard_gpu.py -q -x 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 40 --csv --log-file $OUT/launches_default_$TAG.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-test --no-parity-check > $OUT/ncu_launches_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches_default_$TAG.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
seen={}
for r in rows[1:]:
    seen.setdefault(r[ii],{"k":r[ki][:60]})[r[mi]]=r[vi]
for i,d in list(seen.items())[-16:]:
    print(i, d["k"], "ms", float(d.get("gpu__time_duration.sum",0))/1e6, "dramR GB", float(d.get("dram__bytes_read.sum",0))/1e9, "dramW GB", float(d.get("dram__bytes_write.sum",0))/1e9, "lts GB", float(d.get("lts__t_bytes.sum",0))/1e9)
PY
