"""GPU parity of the FP16 tensor-core filter (option k5_f16: K4h + K5h, wc_search.cu) against the oracle.

EXPERIMENTAL: written at the end of round 1 without GPU time left to run it, so it is opt-in (WC_TEST_F16=1) and the
option is off by default; `bash tools/gpu_f16.sh` runs these tests and the comparison benches."""
import os

import numpy as np
import pytest

import c_oracle
from wisecondor_b200 import _cabi, synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("WC_TEST_F16") != "1", reason="experimental fp16 filter: set WC_TEST_F16=1")]


@pytest.fixture
def f16():
    ctx = _cabi.context(0)

    def setopt(key, value):
        _cabi.check(_cabi.lib().wc_set_option(ctx.handle, key.encode(), float(value)))
    setopt("k5_f16", 1)
    yield setopt
    setopt("k5_f16", 0)
    setopt("k5_sym", 8)


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
    assert st["exhaustive_rows"] == 0


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
    monkeypatch.setenv("WC_K5_F16", "1")
    bins = [int(b) for b in np.array(synth.chrom_bins(250000)) // 2]
    X = synth.corrected_like(bins, 64, seed=31)
    oidx, odist = c_oracle.get_reference_rows(X, bins, 0, X.shape[0], 100)
    for world in (2, 3):
        idx, dist, stats = _emulated_ranks(X, bins, 100, world)
        _assert_same(idx, dist, oidx, odist)
