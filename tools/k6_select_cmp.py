"""K6a comparison: the search on named bench workloads with the warp-per-row select holding a row's entries (k6_select = 0)
and with the streaming histogram select (k6_select = 1); prints K6 times, shortlist sizes, fallback rows, and checks that
both settings return the same table (sha of indexes and distances)."""
import hashlib
import json
import sys

import torch

sys.path.insert(0, ".")
from bench import WORKLOADS  # noqa: E402
from wisecondor_b200 import _cabi, device, synth  # noqa: E402

ctx = _cabi.context(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def opt(key, val):
    _cabi.check(_cabi.lib().wc_set_option(ctx.handle, key.encode(), float(val)))


for name in sys.argv[1:] or ["newref_600x50kb"]:
    binsize, S, k, _ = WORKLOADS[name]
    bins = synth.chrom_bins(binsize)
    if binsize <= 10000:
        X = synth.corrected_like_device(bins, S, seed=4)
    else:
        X = torch.from_numpy(synth.corrected_like(bins, S, seed=4)).cuda()
    n = X.shape[0]
    ref = None
    for select in (0, 1):
        opt("k6_select", select)
        k6, rs, tot = [], [], []
        for it in range(3 if binsize <= 10000 else 5):
            flush.fill_(it)
            torch.cuda.synchronize()
            idx, dist = device.newref_topk(X, bins, 0, n, k)
            torch.cuda.synchronize()
            st = device.last_search_stats(0)
            if it:
                k6.append(st["finalize_ms"]); rs.append(st["finalize_rescore_ms"]); tot.append(st["dist_topk_ms"])
        h = hashlib.sha256(idx.cpu().numpy().tobytes() + dist.cpu().numpy().tobytes()).hexdigest()[:16]
        ref = ref or h
        print(json.dumps({"workload": name, "k6_select": select, "k6_ms": round(min(k6), 3), "rescore_ms": round(min(rs), 3),
                          "select_and_rank_ms": round(min(a - b for a, b in zip(k6, rs)), 3), "k5_ms": round(min(tot), 3),
                          "exhaustive_rows": st["exhaustive_rows"], "live_per_row": round(st["k6_live_entries"] / n, 1),
                          "shortlist_per_row": round(st["k6_shortlisted"] / n, 1), "rows_cta_select": st["k6_rows_cta_select"],
                          "live_max": st["k6_live_max"], "same_table": h == ref}), flush=True)
    del X
