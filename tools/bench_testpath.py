"""Timing of the batched test path (K8 z-scores + K9 segmentation) on a named shape (debug / exploration)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from wisecondor_b200 import device, synth  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
S = int(sys.argv[3]) if len(sys.argv) > 3 else 600
k = 100
bins = synth.chrom_bins(binsize)
n = int(sum(bins))
X = torch.from_numpy(synth.corrected_like(bins, S, seed=4)).cuda()
idx, dist = device.newref_topk(X, bins, 0, n, k)
dist_h = dist.cpu().numpy()
cut = float("inf")
for _ in range(3):
    sel = dist_h[dist_h < cut]
    cut = np.average(sel) + 3 * np.std(sel)
table = device.ReferenceTable(idx.cpu().numpy(), dist_h, bins, cut)
print("N=%d cutoff=%g mean refs/bin=%.1f" % (n, cut, table.count.float().mean().item()))
rng = np.random.default_rng(1)
ldb = device.pad32(B)
T = torch.ones((n, ldb), dtype=torch.float64, device="cuda")
host = 1.0 + rng.normal(0, 0.03, size=(n, B))
for b in range(B):
    a = int(rng.integers(0, n - 400))
    host[a:a + int(rng.integers(20, 400)), b] *= rng.choice([0.9, 1.1])
T[:, :B] = torch.from_numpy(host).cuda()
thr = 5.4
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    z, r, sizes, asdef = device.zscore_batch(T, B, table, thr, 5)
    torch.cuda.synchronize()
    t1 = time.time()
    cwz, cleaned, calls = device.segment_batch(z, sizes, bins, list(range(22)), 25, thr, 3)
    torch.cuda.synchronize()
    t2 = time.time()
    st = device.last_test_stats(0)
    print(json.dumps({"B": B, "zscore_ms": st["zscore_ms"], "segment_ms": st["segment_ms"], "wall_z": t1 - t0, "wall_seg": t2 - t1,
                      "calls": int(len(calls)), "samples_per_s_device": B / ((st["zscore_ms"] + st["segment_ms"]) * 1e-3)}))
