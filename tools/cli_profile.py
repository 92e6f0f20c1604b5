"""Where the wall time of `wisecondor.py newref` goes on BASELINE configs[0] (20 samples x 250 kb): cProfile of the command."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import write_sample_npz
from wisecondor_b200 import synth
binsize = 250000
bins, lam, fac = synth.bin_model(binsize, bin_seed=1)
ref = synth.sample_counts(20, lam, fac, seed=2)
d = tempfile.mkdtemp(prefix="wc_prof_")
for i in range(20):
    write_sample_npz(os.path.join(d, "r%02d.npz" % i), ref[i], bins, binsize)
refs = ["r%02d.npz" % i for i in range(20)]
for rep in range(2):
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "cProfile", "-s", "cumtime", os.path.join(ROOT, "wisecondor.py"), "newref"] + refs + ["g%d.npz" % rep],
                       cwd=d, capture_output=True, text=True)
    print("wall", round(time.time() - t0, 2), "rc", r.returncode)
lines = r.stdout.splitlines()
start = next(i for i, l in enumerate(lines) if "cumulative" in l)
print("\n".join(l[:150] for l in lines[start:start + 45]))
