"""GPU parity of the SHARDED symmetric search (wc_newref_shard_begin / _sweep / _finish): the block pairs of the symmetric
search divided over `world` ranks.  Here the ranks are emulated one after the other on ONE GPU, each with its own library
context, and the two collectives (threshold all-reduce MIN, candidate all-to-all) are done by hand with tensor ops - the
kernels and buffer layouts are exactly those of a multi-GPU run; the collectives themselves are covered by
tests/test_host_cpu.py::test_symmetric_sharded_search_collectives_over_gloo."""
import numpy as np
import pytest

import c_oracle
from wisecondor_b200 import synth

pytestmark = pytest.mark.gpu


def _emulated_ranks(X, bins, k, world):
    import torch
    from wisecondor_b200 import _cabi, device
    dev = torch.device("cuda", 0)
    x = torch.as_tensor(np.ascontiguousarray(X), device=dev)
    n = X.shape[0]
    ctxs = [_cabi.Context(0) for _ in range(world)]
    try:
        dims = [device.shard_dims(n, k, world, r, ctx=ctxs[r]) for r in range(world)]
        rows_per, in_cap, thr_len = dims[0]["rows_per"], dims[0]["in_cap"], dims[0]["thr_len"]
        assert all(d["rows_per"] == rows_per and d["in_cap"] == in_cap for d in dims)
        assert dims[0]["row0"] == 0 and max(d["row1"] for d in dims) == n
        thr = [torch.empty((thr_len,), dtype=torch.int64, device=dev) for _ in range(world)]
        for r in range(world):
            device.shard_begin(x, bins, k, r, world, thr[r], ctx=ctxs[r])
        torch.cuda.synchronize()
        tmin = torch.stack(thr).min(dim=0).values                    # all-reduce(MIN)
        assert int((tmin[:n] < 0).all()) and int((tmin[n:] == 0).all())
        found = []
        for r in range(world):
            in_key = torch.empty((world * rows_per, in_cap), dtype=torch.int64, device=dev)
            in_j = torch.empty((world * rows_per, in_cap), dtype=torch.int32, device=dev)
            in_cnt = torch.empty((world * rows_per,), dtype=torch.int32, device=dev)
            thr_r = tmin.clone()          # stays alive until shard_finish (K6 reads the bins' final thresholds from it)
            device.shard_sweep(thr_r, in_key, in_j, in_cnt, ctx=ctxs[r])
            found.append((in_key, in_j, in_cnt, thr_r))
        torch.cuda.synchronize()
        idx_parts, dist_parts, stats = [], [], []
        for o in range(world):                                       # all-to-all: split o of every rank, in rank order
            sl = slice(o * rows_per, (o + 1) * rows_per)
            recv = [torch.cat([found[s][t][sl] for s in range(world)]).contiguous() for t in range(3)]
            rows = dims[o]["row1"] - dims[o]["row0"]
            idx = torch.empty((rows, k), dtype=torch.int32, device=dev)
            dst = torch.empty((rows, k), dtype=torch.float64, device=dev)
            device.shard_finish(recv[0], recv[1], recv[2], idx, dst, ctx=ctxs[o])
            idx_parts.append(idx.cpu().numpy())
            dist_parts.append(dst.cpu().numpy())
            stats.append((ctxs[o].counter(1), ctxs[o].counter(2), ctxs[o].counter(3)))
        return np.concatenate(idx_parts), np.concatenate(dist_parts), stats
    finally:
        for c in ctxs:
            c.close()


def _assert_same(idx, dist, oidx, odist):
    assert idx.shape == oidx.shape and dist.shape == odist.shape
    bad = np.flatnonzero((idx != oidx).any(axis=1))
    assert bad.size == 0, "index rows differ: %d rows, first %s (%s vs %s)" % (bad.size, bad[:8], idx[bad[0]][:10], oidx[bad[0]][:10])
    assert np.array_equal(dist, odist)


@pytest.mark.parametrize("world,S,k", [(1, 40, 50), (2, 64, 100), (3, 37, 100), (4, 50, 200), (8, 24, 30)])
def test_sharded_symmetric_vs_c_oracle(world, S, k):
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]          # N ~ 5760: 46 blocks of 128 bins
    X = synth.corrected_like(bins, S, seed=20 + S)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], k)
    idx, dist, stats = _emulated_ranks(X, bins, k, world)
    _assert_same(idx, dist, oidx, odist)
    tiles = sum(s[2] for s in stats)
    plain = sum(s[1] for s in stats)
    assert 0.5 < tiles / float(plain) < 0.66, (tiles, plain)                   # every unordered block pair once (+ 1/8)


def test_sharded_symmetric_ties_nan_and_more_ranks_than_blocks():
    rng = np.random.default_rng(9)
    bins = [1500, 700, 1400, 60]
    X = synth.corrected_like(bins, 24, seed=4)
    X[1500:2200] = X[1500]                  # 700 copies of one bin: tie plateaus -> exact fallback rows
    X[5] = X[1500]
    X[2300] = X[3] + rng.normal(0, 1e-9, 24)
    X[10, 3] = np.nan
    X[2500, 0] = np.inf
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    idx, dist, stats = _emulated_ranks(X, bins, 100, 3)
    _assert_same(idx, dist, oidx, odist)
    small = [200, 100, 150]                 # 4 blocks on 8 ranks: most ranks own nothing
    Xs = synth.corrected_like(small, 16, seed=3)
    oidx, odist = c_oracle.get_reference_rows(Xs, small, 0, 450, 20)
    idx, dist, stats = _emulated_ranks(Xs, small, 20, 8)
    _assert_same(idx, dist, oidx, odist)
