"""BASELINE configs[0] end to end through the command lines: newref on 20 synthetic 250 kb samples (refsize 100), then
test of one sample - this build's wisecondor.py on the GPU and the reference's own CLI (oracle/_ref, CPU) on the same
sample files; compares the two result sets and prints wall times.  TEST INFRASTRUCTURE use of oracle/_ref."""
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
from make_golden import write_sample_npz  # noqa: E402
from wisecondor_b200 import synth  # noqa: E402

binsize = 250000
bins, lam, fac = synth.bin_model(binsize, bin_seed=1)
ref = synth.sample_counts(20, lam, fac, seed=2)
test = synth.sample_counts(1, lam, fac, seed=3)
synth.inject_aberration(test[0], bins, 21, 0.0, 1.0, 1.05, seed=4)      # a 5 % whole-chromosome-21 gain
synth.inject_aberration(test[0], bins, 5, 0.3, 0.36, 0.85, seed=5)      # a 15 % deletion, ~40 bins
d = tempfile.mkdtemp(prefix="wc_cfg1_")
for i in range(20):
    write_sample_npz(os.path.join(d, "r%02d.npz" % i), ref[i], bins, binsize)
write_sample_npz(os.path.join(d, "t.npz"), test[0], bins, binsize)
refs = ["r%02d.npz" % i for i in range(20)]
out = {"config": "configs[0]: newref 20 x 250 kb (N raw %d, refsize 100) + test 1 sample" % sum(bins)}


def run(cli, argv):
    t0 = time.time()
    r = subprocess.run([sys.executable, cli] + argv, cwd=d, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("%s %s failed:\n%s\n%s" % (cli, argv[0], r.stdout[-1500:], r.stderr[-1500:]))
    return time.time() - t0


mine = os.path.join(ROOT, "wisecondor.py")
out["b200_newref_wall_s"] = run(mine, ["newref"] + refs + ["gref.npz"])
out["b200_test_wall_s"] = run(mine, ["test", "t.npz", "gout.npz", "gref.npz"])
theirs = os.path.join(ROOT, "oracle", "_ref", "wisecondor.py")
if os.path.isfile(theirs):
    cores = os.cpu_count() or 1
    out["reference_cores"] = cores
    out["reference_newref_wall_s"] = run(theirs, ["newref"] + refs + ["cref.npz", "-cpus", str(cores)])
    out["reference_test_wall_s"] = run(theirs, ["test", "t.npz", "cout.npz", "cref.npz"])
    # cross-check: this build's test on the reference's own reference file vs the reference's result
    run(mine, ["test", "t.npz", "xout.npz", "cref.npz"])
    a = np.load(os.path.join(d, "xout.npz"), allow_pickle=True)
    b = np.load(os.path.join(d, "cout.npz"), allow_pickle=True)
    za, zb = np.concatenate(list(a["results_z"])), np.concatenate(list(b["results_z"]))
    ca, cb = np.asarray(a["results_calls"], dtype=float).reshape(-1, 5), np.asarray(b["results_calls"], dtype=float).reshape(-1, 5)
    out["test_on_reference_npz"] = {
        "max_rel_z": float(np.nanmax(np.abs(za - zb) / np.maximum(np.abs(zb), 1e-300))) if za.shape == zb.shape else None,
        "calls_equal_coordinates": bool(ca.shape == cb.shape and np.array_equal(ca[:, :3], cb[:, :3])),
        "ncalls": int(cb.shape[0]),
        "max_rel_call_z": float(np.max(np.abs(ca[:, 3] - cb[:, 3]) / np.abs(cb[:, 3]))) if ca.shape == cb.shape and len(cb) else 0.0}
    g = np.load(os.path.join(d, "gref.npz"), allow_pickle=True)
    c = np.load(os.path.join(d, "cref.npz"), allow_pickle=True)
    out["newref_vs_reference"] = {
        "mask_equal": bool(np.array_equal(g["mask"], c["mask"])),
        "index_rows_identical": float(np.mean((g["indexes"] == c["indexes"]).all(axis=1))),
        "max_rel_distance": float(np.max(np.abs(g["distances"] - c["distances"]) / np.maximum(c["distances"], 1e-300)))}
print(json.dumps(out))
