#!/bin/bash
# per-kernel durations of the K6 split form (ncu launch list; cold-cache, serialised: shares, not absolutes)
TAG=${1:-r02z3}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:"wc_fin|wc_finalize" -c 12 --csv --log-file $OUT/k6_launches_$TAG.csv python tools/k6_sweep.py newref_600x50kb 4,4,100 > $OUT/k6_launches_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/k6_launches_$TAG.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:40], r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
