"""GPU parity of the SYMMETRIC reference-bin search (the default for whole-matrix calls; option k5_sym: every unordered
pair of bin blocks is contracted once and serves both bins, wc_search.cu) against the oracle and against the plain
search, through the C ABI."""
import numpy as np
import pytest

import c_oracle
from wisecondor_b200 import _cabi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture
def sym():
    """Hands back a function that sets the first-pass fraction (0 = plain search); restores the default afterwards."""
    ctx = _cabi.context(0)

    def set_frac(frac):
        _cabi.check(_cabi.lib().wc_set_option(ctx.handle, b"k5_sym", float(frac)))
    set_frac(8)
    yield set_frac
    set_frac(8)


def _gpu_search(X, bins, k):
    from wisecondor_b200 import device
    idx, dist = device.newref_topk_host(X, bins, 0, X.shape[0], k)
    return idx, dist, device.last_search_stats(0)


def _assert_same(idx, dist, oidx, odist):
    assert idx.shape == oidx.shape and dist.shape == odist.shape
    bad = np.flatnonzero((idx != oidx).any(axis=1))
    assert bad.size == 0, "index rows differ: %d rows, first %s (%s vs %s)" % (bad.size, bad[:8], idx[bad[0]][:10], oidx[bad[0]][:10])
    assert np.array_equal(dist, odist)


@pytest.mark.parametrize("S,k,frac", [(64, 100, 8), (37, 100, 4), (50, 200, 16), (24, 7, 2), (130, 128, 8)])
def test_symmetric_vs_c_oracle(sym, S, k, frac):
    sym(frac)
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]          # N ~ 5760: 46 row blocks
    X = synth.corrected_like(bins, S, seed=5 + S)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], k)
    idx, dist, st = _gpu_search(X, bins, k)
    _assert_same(idx, dist, oidx, odist)
    assert st["launches"] >= 6                                                 # two K5 passes ran


def test_symmetric_equals_plain_search_and_halves_the_tiles(sym):
    bins = synth.chrom_bins(250000)                                            # BASELINE config 2 shape
    X = synth.corrected_like(bins, 600, seed=4)
    idx, dist, st = _gpu_search(X, bins, 100)
    sym(0)
    pidx, pdist, pst = _gpu_search(X, bins, 100)
    _assert_same(idx, dist, pidx, pdist)
    # plain: two fills, K4, K5, K6 (one fused kernel, or two selects + bucket scan + scatter + re-score + rank - once per row
    # range the host call finalises in: option k6_parts, 1..16) (+ pivot select, gather, pass with the tcgen05 filter);
    # symmetric: one K5 launch more
    k6 = pst["launches"] - 4 - (3 if pst["pivots"] else 0)
    assert k6 in (1, 2) or (k6 % 6 == 0 and 1 <= k6 // 6 <= 16)
    assert st["launches"] == pst["launches"] + 1
    assert 0.5 < st["tiles"] / pst["tiles"] < 0.62                             # 1/8 + 7/16 = 0.5625 of the plain tiles
    assert st["tiles_plain"] == pst["tiles"] == pst["tiles_plain"]
    assert st["exhaustive_rows"] == pst["exhaustive_rows"] == 0


def test_symmetric_ties_nan_and_short_candidate_lists(sym):
    rng = np.random.default_rng(9)
    bins = [1500, 700, 1400, 60]
    X = synth.corrected_like(bins, 24, seed=4)
    X[1500:2200] = X[1500]                  # chromosome 2 = 700 copies of one bin: tie plateaus -> exact fallback rows
    X[5] = X[1500]
    X[2300] = X[3] + rng.normal(0, 1e-9, 24)
    X[10, 3] = np.nan
    X[2500, 0] = np.inf
    X[3000] = 1e6                           # distances >= 1e10 are never inserted (wisetools.py:312-314)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    idx, dist, st = _gpu_search(X, bins, 100)
    _assert_same(idx, dist, oidx, odist)
    assert (idx[10] == -1).all() and (idx[2500] == -1).all()
    bins = [3100, 30, 20]                   # rows of chromosome 1 have 50 candidates only: fillers (wisetools.py:305-306)
    X = synth.corrected_like(bins, 16, seed=2)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 60)
    idx, dist, st = _gpu_search(X, bins, 60)
    _assert_same(idx, dist, oidx, odist)
    assert (idx[:3100, 50:] == -1).all()


def test_partial_ranges_and_small_genomes_take_the_plain_path(sym):
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]
    X = synth.corrected_like(bins, 20, seed=8)
    from wisecondor_b200 import device
    idx, dist = device.newref_topk_host(X, bins, 100, 900, 10)
    st = device.last_search_stats(0)
    assert st["launches"] - (3 if st["pivots"] else 0) in (5, 6, 10, 16) and st["tiles"] == st["tiles_plain"]  # one K5 pass (+ the pivot pass)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 100, 900, 10)
    _assert_same(idx, dist, oidx, odist)
    small = [300, 200, 250]
    Xs = synth.corrected_like(small, 20, seed=8)
    idx, dist = device.newref_topk_host(Xs, small, 0, 750, 10)
    assert device.last_search_stats(0)["launches"] in (5, 6, 10, 16)


def test_rescore_by_tma_gather4_returns_the_same_table(sym):
    """Option k6_g4: K6c fetches the candidate rows four per TMA request (tile::gather4) - same table as the bulk-copy form."""
    from wisecondor_b200 import _cabi
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]
    X = synth.corrected_like(bins, 130, seed=21)
    idx, dist, st = _gpu_search(X, bins, 100)
    ctx = _cabi.context(0)
    _cabi.check(_cabi.lib().wc_set_option(ctx.handle, b"k6_g4", 1.0))
    try:
        gidx, gdist, gst = _gpu_search(X, bins, 100)
    finally:
        _cabi.check(_cabi.lib().wc_set_option(ctx.handle, b"k6_g4", 0.0))
    _assert_same(gidx, gdist, idx, dist)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, 256, 100)
    _assert_same(gidx[:256], gdist[:256], oidx, odist)


@pytest.mark.parametrize("S,k", [(130, 100), (50, 200), (64, 20)])
def test_both_selects_return_the_same_table(sym, S, k):
    """Option k6_select: the streaming two-level histogram select (default) and the select that holds a row's entries in
    shared memory (rows with more than 1024 of them: CTA-per-row select) shortlist differently sized supersets of the
    refsize nearest bins; the exact re-score and the rank make the same table of either.  Unpruned rows (small N, large k:
    no published threshold, more entries than a warp holds) included."""
    from wisecondor_b200 import _cabi
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]
    X = synth.corrected_like(bins, S, seed=23 + S)
    idx, dist, st = _gpu_search(X, bins, k)
    assert st["exhaustive_rows"] == 0
    ctx = _cabi.context(0)
    _cabi.check(_cabi.lib().wc_set_option(ctx.handle, b"k6_select", 0.0))
    try:
        hidx, hdist, hst = _gpu_search(X, bins, k)
    finally:
        _cabi.check(_cabi.lib().wc_set_option(ctx.handle, b"k6_select", 1.0))
    _assert_same(hidx, hdist, idx, dist)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, 256, k)
    _assert_same(idx[:256], dist[:256], oidx, odist)
