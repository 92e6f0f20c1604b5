"""profiles/traffic_r03.json from an ncu launch list (gpu__time_duration, dram__bytes_read/write, lts__t_bytes per launch):
DRAM bytes per step of the K5t launches together and of the K6c launch, for bench.py's roofline.traffic.
usage: python tools/traffic_from_csv.py launches.csv workload [launches2.csv workload2 ...] > profiles/traffic_r03.json"""
import csv
import json
import sys

out = {}
for path, workload in zip(sys.argv[1::2], sys.argv[2::2]):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    launches = {}
    for r in rows[1:]:
        launches.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    ls = list(launches.values())
    # the last complete step: from the last wc_prepare_f16_kernel on... take the LAST occurrence of each kernel class
    def last(pred, n=1):
        sel = [l for l in ls if pred(l["k"])]
        return sel[-n:]
    k5 = last(lambda k: "wc_dist_topk_tc_kernel" in k, 3)
    k6c = last(lambda k: "wc_fin_rescore_kernel" in k)
    dram = lambda l: l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    out[workload] = {
        "k5": sum(dram(l) for l in k5), "k5_launches": [{"kernel": l["k"][:60], "ms": l.get("gpu__time_duration.sum", 0.0) / 1e6, "dram_bytes": dram(l),
                                                          "l2_bytes": l.get("lts__t_bytes.sum", 0.0)} for l in k5],
        "k6c": sum(dram(l) for l in k6c), "k6c_ms": sum(l.get("gpu__time_duration.sum", 0.0) for l in k6c) / 1e6,
        "k6c_l2_bytes": sum(l.get("lts__t_bytes.sum", 0.0) for l in k6c),
        "source": path, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, --clock-control none); k5 = pivot + threshold + symmetric pass of one step"}
print(json.dumps(out, indent=1))
