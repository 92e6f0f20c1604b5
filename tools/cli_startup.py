"""Wall time of single `wisecondor.py test` invocations (fresh interpreter each) with the two buffer backends, on a
synthetic 50 kb-shaped reference (N ~ 57k masked bins, refsize 300).  Prints one JSON line."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from wisecondor_b200 import synth                                                          # noqa: E402
from make_golden import write_sample_npz                                                   # noqa: E402


def main():
    rng = np.random.default_rng(5)
    bins = synth.chrom_bins(50000)
    n, k, ncomp = sum(bins), 300, 3
    d = tempfile.mkdtemp(prefix="wc_startup_")
    # a reference npz of the right shape: random other-chromosome indexes, sorted distances, everything unmasked
    starts = np.concatenate(([0], np.cumsum(bins)))
    idx = np.empty((n, k), dtype=np.int32)
    for c in range(len(bins)):
        idx[starts[c]:starts[c + 1]] = rng.integers(0, n - bins[c], size=(bins[c], k))
    dist = np.sort(rng.gamma(4.0, 1e-4, size=(n, k)), axis=1)
    comps = np.linalg.qr(rng.normal(size=(n, ncomp)))[0].T
    np.savez(os.path.join(d, "ref.npz"), arguments={}, runtime={}, binsize=50000, indexes=idx, distances=dist,
             chromosome_sizes=bins, mask=np.ones(n, dtype=bool), masked_sizes=bins, pca_components=comps * 1e-3,
             pca_mean=np.full(n, 1.0 / n))
    write_sample_npz(os.path.join(d, "s.npz"), rng.poisson(120, size=n).astype(np.int32), bins, 50000)
    out = {}
    for backend in ("torch", "native"):
        env = dict(os.environ, WISECONDOR_BACKEND=backend)
        times = []
        for rep in range(3):
            t0 = time.time()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "wisecondor.py"), "test", os.path.join(d, "s.npz"),
                                os.path.join(d, "o_%s.npz" % backend), os.path.join(d, "ref.npz")],
                               capture_output=True, text=True, env=env)
            times.append(time.time() - t0)
            if r.returncode != 0:
                sys.stderr.write(r.stdout[-1500:] + r.stderr[-1500:])
                raise SystemExit(1)
        out[backend + "_s"] = [round(t, 2) for t in times]
    a = np.load(os.path.join(d, "o_torch.npz"), allow_pickle=True)
    b = np.load(os.path.join(d, "o_native.npz"), allow_pickle=True)
    out["identical"] = bool(np.array_equal(np.concatenate(list(a['results_z'])), np.concatenate(list(b['results_z'])), equal_nan=True)
                            and np.array_equal(np.asarray(a['results_calls']), np.asarray(b['results_calls'])))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
