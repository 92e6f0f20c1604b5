"""GPU parity of the reference-bin search (K4+K5+K6) against the oracle, through the C ABI."""
import numpy as np
import pytest

import c_oracle
import wc_oracle
from wisecondor_b200 import synth

pytestmark = pytest.mark.gpu


def _gpu_search(X, bins, r0, r1, k):
    from wisecondor_b200 import device
    return device.newref_topk_host(X, bins, r0, r1, k)


def _assert_same(idx, dist, oidx, odist):
    assert idx.shape == oidx.shape and dist.shape == odist.shape
    bad = np.flatnonzero((idx != oidx).any(axis=1))
    assert bad.size == 0, "index rows differ: %s (first: %s vs %s)" % (bad[:8], idx[bad[0]][:10], oidx[bad[0]][:10])
    assert np.array_equal(dist, odist), "distances not bit-identical: max rel %g" % np.max(
        np.abs(dist - odist) / np.maximum(np.abs(odist), 1e-300))


def test_small_genome_vs_numpy_oracle(small_bins):
    X = synth.corrected_like(small_bins, 40, seed=3)
    sums = np.cumsum(small_bins)
    oidx, odist = wc_oracle.get_reference(X, small_bins, sums, 20, 1, 1)
    idx, dist = _gpu_search(X, small_bins, 0, X.shape[0], 20)
    _assert_same(idx, dist, oidx, odist)


@pytest.mark.parametrize("S,k", [(20, 100), (37, 100), (600, 100), (64, 7), (130, 128), (50, 200)])
def test_medium_vs_c_oracle(S, k):
    bins = [int(b) for b in np.maximum(1, np.array(synth.chrom_bins(250000)) // 4)]
    X = synth.corrected_like(bins, S, seed=11 + S)
    n = X.shape[0]
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, n, k)
    idx, dist = _gpu_search(X, bins, 0, n, k)
    _assert_same(idx, dist, oidx, odist)


def test_parts_concatenate_to_whole(small_bins):
    """README.md:144-145 / wisetools.py:358-361: single- and multi-part runs are identical."""
    from wisecondor_b200 import wisetools
    X = synth.corrected_like(small_bins, 30, seed=5)
    sums = list(np.cumsum(small_bins))
    whole_i, whole_d = wisetools.getReference(X, small_bins, sums, 25, 1, 1)
    for parts in (2, 3, 7):
        pi, pd = [], []
        for p in range(1, parts + 1):
            i, d = wisetools.getReference(X, small_bins, sums, 25, p, parts)
            pi.append(i)
            pd.append(d)
        assert np.array_equal(np.concatenate(pi), whole_i)
        assert np.array_equal(np.concatenate(pd), whole_d)


def test_too_few_candidates_leaves_fillers():
    bins = [30, 4, 3]       # rows of chromosome 1 have only 7 candidates
    X = synth.corrected_like(bins, 16, seed=2)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, 37, 12)
    idx, dist = _gpu_search(X, bins, 0, 37, 12)
    _assert_same(idx, dist, oidx, odist)
    assert (idx[:30, 7:] == -1).all() and (dist[:30, 7:] == 1e10).all()


def test_ties_and_duplicates_fall_back_exactly():
    """Hundreds of identical bins: every distance ties; the order must be by index (wisetools.py:314-320)."""
    bins = [40, 700, 60]
    rng = np.random.default_rng(9)
    X = synth.corrected_like(bins, 24, seed=4)
    X[40:740] = X[40]                       # chromosome 2 = 700 copies of one bin
    X[5] = X[40]                            # and a target on chromosome 1 identical to them (distance 0)
    X[750] = X[3] + rng.normal(0, 1e-9, 24)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    idx, dist = _gpu_search(X, bins, 0, X.shape[0], 100)
    _assert_same(idx, dist, oidx, odist)


def test_nan_and_inf_rows_are_never_selected():
    bins = [50, 60, 40]
    X = synth.corrected_like(bins, 20, seed=6)
    X[10, 3] = np.nan
    X[70, 0] = np.inf
    X[120] = 1e6                            # distances ~ 2e13 >= 1e10: never inserted (wisetools.py:312-314)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 30)
    idx, dist = _gpu_search(X, bins, 0, X.shape[0], 30)
    _assert_same(idx, dist, oidx, odist)
    assert (idx[10] == -1).all() and (idx[70] == -1).all()


def test_row_range_and_empty_range(small_bins):
    X = synth.corrected_like(small_bins, 20, seed=8)
    n = X.shape[0]
    oidx, odist = c_oracle.get_reference_rows(X, small_bins, 100, 333, 10)
    idx, dist = _gpu_search(X, small_bins, 100, 333, 10)
    _assert_same(idx, dist, oidx, odist)
    idx, dist = _gpu_search(X, small_bins, 50, 50, 10)
    assert idx.shape == (0, 10) and dist.shape == (0, 10)
    assert n == sum(small_bins)


def test_config2_shape_sampled_rows():
    """BASELINE config 2 shape (N=11537, S=600, refsize 100): full GPU search, oracle on sampled row ranges."""
    bins = synth.chrom_bins(250000)
    X = synth.corrected_like(bins, 600, seed=4)
    n = X.shape[0]
    idx, dist = _gpu_search(X, bins, 0, n, 100)
    # size-independent properties over every row
    assert (np.diff(dist, axis=1) >= 0).all()
    assert (idx >= 0).all() and (idx < n - np.repeat(bins, bins)[:, None]).all()
    for r0 in (0, 990, 5000, n - 64):
        oidx, odist = c_oracle.get_reference_rows(X, bins, r0, r0 + 64, 100)
        _assert_same(idx[r0:r0 + 64], dist[r0:r0 + 64], oidx, odist)


def test_degenerate_genomes():
    """One chromosome only (no candidate at all), empty chromosomes in the list, refsize 1, fewer bins than a tile."""
    X = synth.corrected_like([25], 10, seed=1)
    idx, dist = _gpu_search(X, [25], 0, 25, 5)
    assert (idx == -1).all() and (dist == 1e10).all()                 # wisetools.py:305-306 fillers only
    bins = [0, 9, 0, 0, 14, 3, 0]
    X = synth.corrected_like(bins, 7, seed=2)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, 26, 1)
    idx, dist = _gpu_search(X, bins, 0, 26, 1)
    _assert_same(idx, dist, oidx, odist)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 3, 20, 6)
    idx, dist = _gpu_search(X, bins, 3, 20, 6)
    _assert_same(idx, dist, oidx, odist)


def test_single_sample_and_wide_refsize():
    bins = [40, 35, 30, 500]
    X = synth.corrected_like(bins, 1, seed=3)                          # S = 1: distances are single squares
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, 605, 64)
    idx, dist = _gpu_search(X, bins, 0, 605, 64)
    _assert_same(idx, dist, oidx, odist)
    X = synth.corrected_like(bins, 33, seed=4)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, 605, 384)     # the largest refsize the kernels accept
    idx, dist = _gpu_search(X, bins, 0, 605, 384)
    _assert_same(idx, dist, oidx, odist)
    with pytest.raises(Exception):
        _gpu_search(X, bins, 0, 605, 385)
    with pytest.raises(Exception):
        _gpu_search(X, [40, 35], 0, 75, 10)                            # chromosome sizes do not add up to N


def test_config3_shape_parts_and_sampled_rows():
    """BASELINE configs[2] shape (N = 57 633, S = 600, refsize 100), device-resident: the 8 getPart shards concatenate
    to the single-part result bit for bit, rows are sorted with valid indices, sampled rows equal the C oracle."""
    import torch
    from wisecondor_b200 import device, shard
    bins = synth.chrom_bins(50000)
    Xh = synth.corrected_like(bins, 600, seed=4)
    X = torch.from_numpy(Xh).cuda()
    n = Xh.shape[0]
    idx, dist = device.newref_topk(X, bins, 0, n, 100)
    idx_h, dist_h = idx.cpu().numpy(), dist.cpu().numpy()
    parts_i, parts_d = [], []
    for r in range(8):
        a, b = shard.row_shard(r, 8, n)
        pi, pd = device.newref_topk(X, bins, a, b, 100)
        parts_i.append(pi.cpu().numpy())
        parts_d.append(pd.cpu().numpy())
    assert np.array_equal(np.concatenate(parts_i), idx_h) and np.array_equal(np.concatenate(parts_d), dist_h)
    assert (np.diff(dist_h, axis=1) >= 0).all()
    assert (idx_h >= 0).all() and (idx_h < n - np.repeat(bins, bins)[:, None]).all()
    for r0 in (0, 4970, 30000, n - 32):                  # 4970: straddles the chromosome 1 / 2 boundary
        oidx, odist = c_oracle.get_reference_rows(Xh, bins, r0, r0 + 32, 100)
        _assert_same(idx_h[r0:r0 + 32], dist_h[r0:r0 + 32], oidx, odist)
