"""K6 tuning aid: the search on a named bench workload for several (k6_split, k6_warps, k6_prod, k6_chunk) settings;
prints K6 / re-score times and checks that every setting returns the same table (sha of indexes and distances)."""
import hashlib
import json
import sys

import torch

sys.path.insert(0, ".")
from bench import WORKLOADS  # noqa: E402
from wisecondor_b200 import _cabi, device, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "newref_600x50kb"
binsize, S, k, _ = WORKLOADS[name]
bins = synth.chrom_bins(binsize)
X = torch.from_numpy(synth.corrected_like(bins, S, seed=4)).cuda()
n = X.shape[0]
ctx = _cabi.context(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def opt(key, val):
    _cabi.check(_cabi.lib().wc_set_option(ctx.handle, key.encode(), float(val)))


combos = [(0, 0, 0, 0, 0)]
for arg in sys.argv[2:]:
    f = [int(v) for v in arg.split(",")]
    combos.append((1, f[0], f[1], f[2], f[3] if len(f) > 3 else 1))          # consumer warps, producer warps each, chunk, gather4
ref = None
for split, w, p, c, g4 in combos:
    opt("k6_split", split); opt("k6_warps", w); opt("k6_prod", p); opt("k6_chunk", c); opt("k6_g4", g4)
    k6, rs, tot = [], [], []
    for it in range(4):
        flush.fill_(it)
        torch.cuda.synchronize()
        idx, dist = device.newref_topk(X, bins, 0, n, k)
        torch.cuda.synchronize()
        st = device.last_search_stats(0)
        if it:
            k6.append(st["finalize_ms"]); rs.append(st["finalize_rescore_ms"]); tot.append(st["dist_topk_ms"])
    h = hashlib.sha256(idx.cpu().numpy().tobytes() + dist.cpu().numpy().tobytes()).hexdigest()[:16]
    ref = ref or h
    print(json.dumps({"workload": name, "split": split, "warps": w, "prod": p, "chunk": c, "gather4": g4, "k6_ms": round(min(k6), 3),
                      "rescore_ms": round(min(rs), 3), "k5_ms": round(min(tot), 3), "exhaustive_rows": st["exhaustive_rows"], "live_per_row": round(st["k6_live_entries"] / n, 1), "shortlist_per_row": round(st["k6_shortlisted"] / n, 1), "rows_cta_select": st["k6_rows_cta_select"], "live_max": st["k6_live_max"],
                      "same_table_as_fused": h == ref}), flush=True)
