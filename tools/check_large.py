"""Parity at BASELINE configs[4] size (2000 samples x 10 kb bins, N = 288113): the GPU search on sampled row ranges of
the device-generated matrix against the C restatement of the oracle (TEST INFRASTRUCTURE use of oracle/)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import c_oracle  # noqa: E402
from wisecondor_b200 import device, synth  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rows_per_range = int(sys.argv[3]) if len(sys.argv) > 3 else 32
bins = synth.chrom_bins(binsize)
n = int(sum(bins))
X = synth.corrected_like_device(bins, S, seed=4, device=torch.device("cuda", 0))
Xh = X.cpu().numpy()
out = {"bins": n, "samples": S, "ranges": []}
for r0 in (0, 24900, n // 2, n - rows_per_range):
    t0 = time.time()
    idx, dist = device.newref_topk(X, bins, r0, r0 + rows_per_range, 100)
    torch.cuda.synchronize()
    t1 = time.time()
    oidx, odist = c_oracle.get_reference_rows(Xh, bins, r0, r0 + rows_per_range, 100)
    t2 = time.time()
    ok = bool(np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(dist.cpu().numpy(), odist))
    out["ranges"].append({"row0": r0, "rows": rows_per_range, "identical": ok, "gpu_s": t1 - t0, "oracle_s": t2 - t1})
print(json.dumps(out))
sys.exit(0 if all(r["identical"] for r in out["ranges"]) else 1)
