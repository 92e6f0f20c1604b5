"""Per-CTA cycle breakdown of the distance/top-k kernel (K5) on a named bench workload (debug aid)."""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import WORKLOADS  # noqa: E402
from wisecondor_b200 import _cabi, device, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "newref_600x250kb"
binsize, S, k, _ = WORKLOADS[name]
bins = synth.chrom_bins(binsize)
X = torch.from_numpy(synth.corrected_like(bins, S, seed=4)).cuda()
n = X.shape[0]
ctx = _cabi.context(0)
_cabi.lib().wc_debug_profile(ctx.handle, 1, None, 0)
if len(sys.argv) > 2:
    _cabi.check(_cabi.lib().wc_set_option(ctx.handle, b"k5_lag", float(sys.argv[2])))
for kv in [x for x in sys.argv[3:] if '=' in x]:
    key, val = kv.split("=")
    _cabi.check(_cabi.lib().wc_set_option(ctx.handle, key.encode(), float(val)))
for _ in range(2):
    device.newref_topk(X, bins, 0, n, k)
buf = np.zeros((320, 8), dtype=np.int64)   # 148 CTAs x 8 counters, then 512 timeline stamps of CTA 0
g = _cabi.lib().wc_debug_profile(ctx.handle, 1, buf.ctypes.data_as(ctypes.c_void_p), 320)
p = buf[:g].astype(float)
st = device.last_search_stats(0)
out = {"workload": name, "lag": (sys.argv[2] if len(sys.argv) > 2 else "default"), "ctas": g, "k5_ms": st["dist_topk_ms"], "k5_first_pass_ms": st["dist_topk_first_pass_ms"],
       "tiles": st["tiles"], "exhaustive_rows": st["exhaustive_rows"], "finalize_ms": st["finalize_ms"],
       "cycles_total_max": p[:, 0].max(), "cycles_total_mean": p[:, 0].mean(), "cycles_total_min": p[:, 0].min(),
       "wait_tma_frac": (p[:, 1] / p[:, 0]).mean(), "epilogue_frac": (p[:, 2] / p[:, 0]).mean(),
       "prune_frac": (p[:, 3] / p[:, 0]).mean(), "tiles_per_cta": p[:, 4].mean(),
       "cycles_per_tile": (p[:, 0] / p[:, 4]).mean(), "col_side_frac": (p[:, 5] / p[:, 0]).mean(),
       "survivors_frac": (p[:, 6] / p[:, 0]).mean(), "flush_frac": (p[:, 7] / p[:, 0]).mean(),
       "ideal_cycles_per_tile": 128 * 128 * S / 64.0}
print(json.dumps(out))
if g + 64 <= 320 and "-t" in sys.argv:
    tl = buf.reshape(-1)[g * 8: g * 8 + 512].reshape(2, 64, 4).astype(np.int64)
    for i in list(range(4, 8)) + list(range(36, 50)):
        row = []
        for w in (0, 1):
            t = tl[w, i]
            row.append("warp%d main %7d epi %6d prune %6d" % (w * 4, t[1] - t[0], t[2] - t[1], t[3] - t[2]))
        print("tile %2d  " % i + " | ".join(row))
