"""CPU: the oracle against the reference itself (oracle/_ref, generated from /root/reference by
oracle/make_ref.py) on random cases.  Skipped where oracle/_ref is absent; tests/test_oracle_golden.py holds the
committed pins."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

import c_oracle
import wc_oracle

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "wisetools.py")),
                                reason="oracle/_ref not generated (needs /root/reference)")


@pytest.fixture(scope="module")
def ref_wt():
    sys.path.insert(0, REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import wisetools as m
        yield m
    finally:
        sys.path.remove(REF)


def _quiet(f, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a)


@pytest.mark.parametrize("seed", range(4))
def test_search_random(ref_wt, seed):
    rng = np.random.default_rng(seed)
    bins = [int(b) for b in rng.integers(1, 25, size=int(rng.integers(2, 7)))]
    n, S, k = sum(bins), int(rng.integers(1, 40)), int(rng.integers(1, 30))
    X = 1.0 + rng.normal(0, 0.05, size=(n, S))
    X[rng.integers(0, n)] = X[rng.integers(0, n)]
    sums = list(np.cumsum(bins))
    ridx, rdst = _quiet(ref_wt.getReference, np.asfortranarray(X), bins, sums, k, 1, 1)
    oidx, odst = wc_oracle.get_reference(X, bins, sums, k, 1, 1)
    cidx, cdst = c_oracle.get_reference_rows(X, bins, 0, n, k)
    assert np.array_equal(ridx, oidx) and np.array_equal(rdst, odst)
    assert np.array_equal(ridx, cidx) and np.array_equal(rdst, cdst)


@pytest.mark.parametrize("seed", range(3))
def test_zscores_random(ref_wt, seed):
    rng = np.random.default_rng(100 + seed)
    bins = [int(b) for b in rng.integers(5, 30, size=4)]
    n = sum(bins)
    sums = list(np.cumsum(bins))
    X = 1.0 + rng.normal(0, 0.05, size=(n, 10))
    idx, dst = wc_oracle.get_reference(X, bins, sums, 9, 1, 1)
    cutoff, _ = ref_wt.getOptimalCutoff(dst, 3)
    assert cutoff == wc_oracle.get_optimal_cutoff(dst, 3)
    test = 1.0 + rng.normal(0, 0.03, size=n)
    test[3:9] *= 1.3
    want = _quiet(ref_wt.repeatTest, np.copy(test), idx, dst, bins, sums, cutoff, 2.5, 4)
    got = wc_oracle.repeat_test(np.copy(test), idx, dst, bins, sums, cutoff, 2.5, 4)
    for a, b in zip(want[:3], got[:3]):
        assert np.array_equal(a, b, equal_nan=True)
    assert want[3] == got[3]


@pytest.mark.parametrize("seed", range(4))
def test_segmentation_random(ref_wt, seed):
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(1, 160))
    z = rng.normal(0, 1, size=n)
    if n > 20:
        a = int(rng.integers(0, n - 10))
        z[a:a + int(rng.integers(2, 10))] += rng.choice([-2.0, 2.0])
    tri = ref_wt.fillTri(z)
    want = tri.segmentTri(3.0, 3)
    cw, got = wc_oracle.segment_region(z, 3.0, 3)
    assert cw == tri.getValue(0, n - 1)
    assert got == [(float(v), (int(x), int(y))) for v, (x, y) in want] or got == want
