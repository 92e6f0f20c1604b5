"""GPU parity of the FP16 tensor-core filters of the reference-bin search against the oracle: option k5_f16 = 1 (K4h +
K5h, mma.sync, wc_search_f16.cuh) and k5_f16 = 2 (K4h + K5t, tcgen05 with TMEM accumulators, wc_search_tc.cuh).  The
filter only shortlists; K6's exact fp64 re-score decides, so every result must be bit-identical to the oracle's."""
import ctypes

import numpy as np
import pytest

import c_oracle
from wisecondor_b200 import _cabi, synth

pytestmark = [pytest.mark.gpu]


@pytest.fixture(params=[1, 2], ids=["mma_sync", "tcgen05"])
def f16(request):
    ctx = _cabi.context(0)

    def setopt(key, value):
        _cabi.check(_cabi.lib().wc_set_option(ctx.handle, key.encode(), float(value)))
    setopt.mode = request.param
    setopt("k5_f16", request.param)
    yield setopt
    setopt("k5_f16", DEFAULT_FILTER)
    setopt("k5_sym", 8)


DEFAULT_FILTER = int(__import__("os").environ.get("WC_K5_F16", "2"))      # the context default: tcgen05 filter


def _search(X, bins, r0, r1, k):
    from wisecondor_b200 import device
    idx, dist = device.newref_topk_host(X, bins, r0, r1, k)
    return idx, dist, device.last_search_stats(0)


def _assert_same(idx, dist, oidx, odist):
    assert idx.shape == oidx.shape and dist.shape == odist.shape
    bad = np.flatnonzero((idx != oidx).any(axis=1))
    assert bad.size == 0, "index rows differ: %d rows, first %s (%s vs %s)" % (bad.size, bad[:8], idx[bad[0]][:10], oidx[bad[0]][:10])
    assert np.array_equal(dist, odist)


@pytest.mark.parametrize("S,k", [(64, 100), (37, 100), (600, 100), (50, 200), (1, 20), (130, 128)])
def test_f16_filter_small_genome_plain_search(f16, S, k):
    bins = [int(b) for b in np.maximum(1, np.array(synth.chrom_bins(250000)) // 4)]       # N ~ 2880: plain search
    X = synth.corrected_like(bins, S, seed=11 + S)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], k)
    idx, dist, st = _search(X, bins, 0, X.shape[0], k)
    _assert_same(idx, dist, oidx, odist)
    if S > 1:          # one sample: distances collapse onto each other within the fp16 margin, the exact fallback takes the rows
        assert st["exhaustive_rows"] == 0


def test_tc_filter_distances_stay_inside_the_margin():
    """Every filter distance the tcgen05 kernel computes (debug dump) against the fp64 distance: the error must stay below
    the a-priori bound eps * (n_i + n_j) the margins are built on (wc_search.cu: eps16).  Also proves the operand
    descriptors / TMEM layout: a transposed or mis-swizzled tile would be off by O(1), not O(1e-4)."""
    import torch
    ctx = _cabi.context(0)
    L = _cabi.lib()
    bins = [int(b) for b in np.maximum(1, np.array(synth.chrom_bins(250000)) // 4)]
    for S in (64, 600):
        X = synth.corrected_like(bins, S, seed=3 + S)
        n = X.shape[0]
        out = torch.full((n, n), float("nan"), dtype=torch.float32, device="cuda:0")
        _cabi.check(L.wc_set_option(ctx.handle, b"k5_f16", 2.0))
        _cabi.check(L.wc_set_option(ctx.handle, b"k5_sym", 0.0))
        _cabi.check(L.wc_debug_filter_scores(ctx.handle, ctypes.c_void_p(out.data_ptr()), n))
        try:
            from wisecondor_b200 import device
            device.newref_topk(torch.as_tensor(X, device="cuda:0"), bins, 0, n, 50)
            torch.cuda.synchronize()
        finally:
            _cabi.check(L.wc_debug_filter_scores(ctx.handle, None, 0))
            _cabi.check(L.wc_set_option(ctx.handle, b"k5_f16", float(DEFAULT_FILTER)))
            _cabi.check(L.wc_set_option(ctx.handle, b"k5_sym", 8.0))
        got = out.cpu().numpy().astype(np.float64)
        xc = X - 1.0
        nrm = (xc * xc).sum(axis=1)
        exact = nrm[:, None] + nrm[None, :] - 2.0 * (xc @ xc.T)
        computed = ~np.isnan(got)
        # every pair of bins on different chromosomes must have been computed
        chrom = np.repeat(np.arange(len(bins)), bins)
        other = chrom[:, None] != chrom[None, :]
        assert computed[other].all(), "tiles missing: %d of %d" % ((~computed[other]).sum(), other.sum())
        err = np.abs(got - exact)[computed] / (nrm[:, None] + nrm[None, :])[computed]
        ldh = (S + 63) // 64 * 64
        eps16 = 2.0 ** -10 * (1 + 2.0 ** -11) + ldh * 2.0 ** -23 + 2.0 ** -21
        print("S=%d: max filter error %.3g of (n_i + n_j), a-priori bound %.3g" % (S, err.max(), eps16))
        assert err.max() <= 0.5 * eps16


@pytest.mark.parametrize("S,k", [(64, 100), (37, 30), (50, 200)])
def test_f16_filter_symmetric_search(f16, S, k):
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]                       # N ~ 5760: symmetric search
    X = synth.corrected_like(bins, S, seed=5 + S)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], k)
    idx, dist, st = _search(X, bins, 0, X.shape[0], k)
    _assert_same(idx, dist, oidx, odist)
    assert st["launches"] >= 6
    idx, dist, st = _search(X, bins, 1000, 1900, k)                                        # a row range: plain search
    _assert_same(idx, dist, oidx[1000:1900], odist[1000:1900])


def test_f16_filter_edge_cases(f16):
    rng = np.random.default_rng(9)
    bins = [1500, 700, 1400, 60]
    X = synth.corrected_like(bins, 24, seed=4)
    X[1500:2200] = X[1500]                  # tie plateaus -> exact fallback rows
    X[5] = X[1500]
    X[2300] = X[3] + rng.normal(0, 1e-9, 24)
    X[10, 3] = np.nan
    X[2500, 0] = np.inf
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    idx, dist, st = _search(X, bins, 0, X.shape[0], 100)
    _assert_same(idx, dist, oidx, odist)
    X[3000] = 1e6                           # outside fp16's range: the call must fall back to the fp64 filter
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    idx, dist, st = _search(X, bins, 0, X.shape[0], 100)
    _assert_same(idx, dist, oidx, odist)
    Y = synth.corrected_like(bins, 24, seed=5) * 1e-7 + 1.0     # |x - 1| ~ 1e-8: fp16 subnormals / flush to zero
    oidx, odist = c_oracle.get_reference_rows(Y, bins, 0, Y.shape[0], 50)
    idx, dist, st = _search(Y, bins, 0, Y.shape[0], 50)
    _assert_same(idx, dist, oidx, odist)


def test_f16_filter_config2_shape_equals_fp64_filter(f16):
    bins = synth.chrom_bins(250000)
    X = synth.corrected_like(bins, 600, seed=4)
    idx, dist, st = _search(X, bins, 0, X.shape[0], 100)
    f16("k5_f16", 0)
    pidx, pdist, pst = _search(X, bins, 0, X.shape[0], 100)
    _assert_same(idx, dist, pidx, pdist)
    assert st["exhaustive_rows"] == 0


def test_f16_filter_sharded_symmetric_search(monkeypatch):
    """The sharded symmetric search with the fp16 filter, ranks emulated on one GPU (contexts read WC_K5_F16 at creation)."""
    from test_search_shard_gpu import _emulated_ranks
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]
    X = synth.corrected_like(bins, 64, seed=31)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    for mode in ("1", "2"):
        monkeypatch.setenv("WC_K5_F16", mode)
        for world in (2, 3):
            idx, dist, stats = _emulated_ranks(X, bins, 100, world)
            _assert_same(idx, dist, oidx, odist)


def test_pivot_selection_is_the_r_smallest_norms_sorted_by_bin():
    """K5t's pivot pass starts from the R bins of smallest norm (ties by bin index), sorted by bin (wc_pivot_select_kernel)."""
    import torch
    from wisecondor_b200 import _cabi
    rng = np.random.default_rng(3)
    ctx = _cabi.context(0)
    for n, r, ties in ((57633, 512, False), (4100, 512, True), (288113, 512, False), (2048, 512, True)):
        v = rng.random(n).astype(np.float32) * 50.0
        if ties:
            v = np.round(v)                       # massive ties, also at the R-th place
        v[rng.integers(0, n, 5)] = np.inf         # padding-like entries never come first
        order = np.lexsort((np.arange(n), v))[:r]
        want = np.sort(order)
        vd = torch.from_numpy(v).cuda()
        ids = torch.empty(r, dtype=torch.int32, device="cuda")
        _cabi.check(_cabi.lib().wc_debug_pivot_select(ctx.handle, vd.data_ptr(), n, r, ids.data_ptr()))
        assert np.array_equal(ids.cpu().numpy(), want)
