"""BASELINE configs[3]: batched test + z-score segmentation of 10 000 synthetic samples at 50 kb bins, sample-sharded over the
GPUs of one box (no communication in the data path).  Run under torchrun (or plainly for one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/test_10k.py --samples 10000

Every rank builds the same reference (the newref search of a seeded synthetic 600-sample matrix on its own GPU), takes its
partition.sample_shard of the samples (sample i is generated from seed 1000 + i, so any rank can regenerate any sample) and
pushes them through prep (K7) -> z-scores (K8) -> segmentation (K9) in device batches.  Checks: (a) every rank re-tests
`--check` samples of its NEIGHBOUR's shard and the calls / chromosome-wide z / sigma averages must be bit-identical - the result
of a sample cannot depend on the shard it lands in; (b) rank 0 runs the numpy restatement of the oracle on `--oracle` samples
(TEST INFRASTRUCTURE use of oracle/).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wisecondor_b200 import device, partition, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=10000)
ap.add_argument("--binsize", type=int, default=50000)
ap.add_argument("--refsamples", type=int, default=600)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--check", type=int, default=32)
ap.add_argument("--oracle", type=int, default=1)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

bins = synth.chrom_bins(args.binsize)
n = int(sum(bins))
k = 100
thr = 5.4            # norm.ppf(1 - 1/(57633 * 0.5 * 1000))

# ---- the reference: built by the newref search on every rank (seeded: identical everywhere) ----
Xh = synth.corrected_like(bins, args.refsamples, seed=4)
X = torch.from_numpy(Xh).to(dev)
idx_d, dist_d = device.newref_topk(X, bins, 0, n, k)
idx_h, dist_h = idx_d.cpu().numpy(), dist_d.cpu().numpy()
cut = float("inf")
for _ in range(3):                                  # getOptimalCutoff (wisetools.py:328-336)
    sel = dist_h[dist_h < cut]
    cut = np.average(sel) + 3 * np.std(sel)
table = device.ReferenceTable(idx_h, dist_h, bins, cut, device=local)
masked_raw = torch.arange(n, dtype=torch.int32, device=dev)
mean = X.mean(dim=1) / X.mean(dim=1).sum()
comps = torch.zeros((3, n), dtype=torch.float64, device=dev)
comps[0, 0::3] = 1.0
comps[1, 1::3] = 1.0
comps[2, 2::3] = 1.0
comps /= comps.norm(dim=1, keepdim=True)
del X

lam = np.random.default_rng(99).gamma(20.0, 8.7, size=n)


def sample_counts(i):
    """Raw count vector of sample i (any rank can build any sample): Poisson around the shared bin profile, one aberration."""
    rng = np.random.default_rng(1000 + i)
    c = rng.poisson(lam).astype(np.int32)
    a = int(rng.integers(0, n - 500))
    w = int(rng.integers(20, 400))
    c[a:a + w] = (c[a:a + w] * rng.choice([0.8, 1.25])).astype(np.int32)
    return c


def run_batch(counts_pinned, nb, full_copy, host):
    """One device batch from pinned host counts; returns (per-sample results, device ms of prep / z / seg)."""
    counts = counts_pinned[:nb].to(dev, non_blocking=True)
    T = device.test_prep(counts, masked_raw, mean, comps)
    z, r, sizes, asdef = device.zscore_batch(T, nb, table, thr, 5)
    cwz, cleaned, calls = device.segment_batch(z, sizes, bins, list(range(22)), 25, thr, 3)
    st = device.last_test_stats(local)
    if full_copy:           # the per-bin vectors a result npz holds (results_z, results_r; refsizes decide the kept bins)
        host[0][:nb].copy_(z, non_blocking=True)
        host[1][:nb].copy_(r, non_blocking=True)
        host[2][:nb].copy_(sizes, non_blocking=True)
    cwz_h = cwz.cpu().numpy()
    asdef_h = asdef.cpu().numpy()
    torch.cuda.synchronize(dev)
    return calls, cwz_h, asdef_h, (st["prep_ms"], st["zscore_ms"], st["segment_ms"])


a, b = partition.sample_shard(rank, world, args.samples)
mine = b - a
B = args.batch
t_gen0 = time.time()
counts_all = torch.empty((mine, n), dtype=torch.int32).pin_memory()
ca = counts_all.numpy()
for i in range(mine):
    ca[i] = sample_counts(a + i)
t_gen = time.time() - t_gen0
host = [torch.empty((B, n), dtype=torch.float64).pin_memory(), torch.empty((B, n), dtype=torch.float64).pin_memory(),
        torch.empty((B, n), dtype=torch.int32).pin_memory()]

run_batch(counts_all, min(B, mine), False, host)          # warm-up (allocations)
results = {}
timings = {}
for mode in ("compact", "full"):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.time()
    dev_ms = np.zeros(3)
    ncalls = 0
    calls_by_sample, cwz_by_sample, asdef_by_sample = {}, {}, {}
    for first in range(0, mine, B):
        nb = min(B, mine - first)
        calls, cwz_h, asdef_h, ms = run_batch(counts_all[first:first + nb], nb, mode == "full", host)
        dev_ms += ms
        ncalls += len(calls)
        if mode == "compact":
            for s in range(nb):
                cs = calls[calls["sample"] == s]
                calls_by_sample[a + first + s] = [(int(c["chrom"]), int(c["x"]), int(c["y"]), float(c["z"])) for c in cs]
                cwz_by_sample[a + first + s] = cwz_h[s].tobytes()
                asdef_by_sample[a + first + s] = float(asdef_h[s])
    torch.cuda.synchronize(dev)
    timings[mode] = (time.time() - t0, dev_ms.copy(), ncalls)
    if mode == "compact":
        results = (calls_by_sample, cwz_by_sample, asdef_by_sample)

# ---- check (a): a neighbour's samples on this GPU ----
ok_shard, checked = True, 0
if world > 1 and args.check > 0:
    nr = (rank + 1) % world
    na, nbnd = partition.sample_shard(nr, world, args.samples)
    take = min(args.check, nbnd - na)
    cc = torch.empty((take, n), dtype=torch.int32).pin_memory()
    for i in range(take):
        cc.numpy()[i] = sample_counts(na + i)
    calls, cwz_h, asdef_h, _ = run_batch(cc, take, False, host)
    mine_view = {na + s: ([(int(c["chrom"]), int(c["x"]), int(c["y"]), float(c["z"])) for c in calls[calls["sample"] == s]],
                          cwz_h[s].tobytes(), float(asdef_h[s])) for s in range(take)}
    gathered = [None] * world
    dist.all_gather_object(gathered, {i: (results[0][i], results[1][i], results[2][i]) for i in range(a, min(b, a + args.check))})
    theirs = gathered[nr]
    for i, v in mine_view.items():
        checked += 1
        if theirs.get(i) != v:
            ok_shard = False
# ---- check (b): the oracle on rank 0's first samples ----
ok_oracle, oracle_s = None, 0.0
if rank == 0 and args.oracle > 0:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wc_oracle
    sums = [int(v) for v in np.cumsum(bins)]
    ok_oracle = True
    t0 = time.time()
    for i in range(min(args.oracle, mine)):
        # the prepared vector of sample a+i exactly as the device sees it
        T = device.test_prep(counts_all[i:i + 1].to(dev), masked_raw, mean, comps)
        test = T[:, 0].cpu().numpy().copy()
        oz, orr, osz, osd = wc_oracle.repeat_test(test, idx_h, dist_h, bins, sums, cut, thr, 5)
        want = []
        for c in range(22):
            lo, hi = sums[c] - bins[c], sums[c]
            ocw, osegs = wc_oracle.segment_region(oz[lo:hi][osz[lo:hi] >= 25], thr, 3)
            want += [(c, int(x), int(y), float(v)) for v, (x, y) in osegs]
        got = sorted(results[0][a + i])
        if sorted(want) != got or osd != results[2][a + i]:
            ok_oracle = False
    oracle_s = time.time() - t0

t = torch.tensor([timings["compact"][0], timings["full"][0], timings["compact"][1].sum(), float(timings["compact"][2]),
                  1.0 if ok_shard else 0.0, float(checked), t_gen], dtype=torch.float64, device=dev)
tmax = t.clone()
tsum = t.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    tmin = t.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
else:
    tmin = t
if rank == 0:
    line = {
        "workload": "test_%dx%dkb" % (args.samples, args.binsize // 1000), "baseline_config": "configs[3]", "n_gpus": world,
        "samples": args.samples, "bins": n, "refsize": k, "repeats": 5, "batch": B,
        "parallelism": "samples sharded over %d GPU(s) (partition.sample_shard), no communication" % world,
        "device_samples_per_s": args.samples / (float(tmax[2]) * 1e-3),
        "device_ms_max_over_ranks": float(tmax[2]),
        "phases_ms_rank0": {"prep": float(timings["compact"][1][0]), "zscore": float(timings["compact"][1][1]), "segment": float(timings["compact"][1][2])},
        "e2e_compact": {"samples_per_s": args.samples / float(tmax[0]), "wall_s": float(tmax[0]),
                        "h2d_bytes_per_sample": n * 4, "d2h_bytes_per_sample": 22 * 8 + 8,
                        "what": "pinned host counts -> device; calls, chromosome-wide z and sigma averages back (what a report needs)"},
        "e2e_full": {"samples_per_s": args.samples / float(tmax[1]), "wall_s": float(tmax[1]),
                     "h2d_bytes_per_sample": n * 4, "d2h_bytes_per_sample": n * 20 + 22 * 8 + 8,
                     "what": "also the per-bin z, ratio and refsize vectors of every sample back to pinned host memory (a result npz)"},
        "calls_total": int(tsum[3]),
        "sharding_check": {"samples_retested_on_a_neighbour_gpu": int(tsum[5]), "identical": bool(float(tmin[4]) == 1.0)} if world > 1 else None,
        "oracle_check": {"samples": min(args.oracle, mine), "identical": ok_oracle, "oracle_s": oracle_s},
        "host_generation_s_max": float(tmax[6]),
    }
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
