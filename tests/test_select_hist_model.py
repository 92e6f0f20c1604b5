"""CPU: the arithmetic behind the streaming two-level histogram select (wc_search_fin.cuh: wc_fin_select_hist_kernel).

The select never sorts a row's candidate entries.  It histograms their filter distances over 1024 equal buckets of
[d0, dmax] with the bucket index computed in fp32, takes the upper edge of the bucket holding the k-th smallest as a bound
v* of the k-th smallest FROM ABOVE, and resolves that bucket 1024 times finer in a second pass.  The final table is exact
whatever bound is used, provided it is never BELOW the k-th smallest: a bound below it would shortlist fewer than the
refsize nearest candidates.  The kernel covers the fp32 roundings of the bucket index (conversion, multiply) and of the
edge's division by slack factors (1 + 8 * 2^-23 on the coarse edge, 1e-3 of a sub-bucket on the fine one); this file
replays both formulas in numpy with the same operand types on random and adversarial (entries sitting on bucket edges)
data and checks the property."""
import numpy as np
import pytest

BINS = 1024
F32 = np.float32


def _bucket(d, d0, scale):
    """wc_fin_select_hist_kernel: bucket()."""
    with np.errstate(invalid="ignore"):
        x = (d - d0).astype(F32) * scale                               # (float)(d - d0) * scale
        b = np.where(x < F32(BINS), np.where(x >= 0, np.nan_to_num(x, nan=0.0, posinf=0.0).astype(np.int64), 0), BINS - 1)
    return np.where(np.isnan(x), BINS - 1, b)


def _kth_bucket(hist, kk):
    """The bucket of the kk-th smallest counted entry and the number of entries before it (kth_bucket)."""
    c = np.cumsum(hist)
    if c[-1] < kk:
        return -1, 0
    b = int(np.searchsorted(c, kk))
    return b, int(c[b] - hist[b])


def _select_bounds(d, k, d0, dmax):
    """(coarse bound v*, refined bound) exactly as the kernel forms them; None where it keeps dmax."""
    span = dmax - d0
    scale = F32(BINS) / F32(span) if span > 1e-5 * abs(dmax) else F32(0)
    bk = _bucket(d, d0, scale)
    bstar, below = _kth_bucket(np.bincount(bk, minlength=BINS), k)
    if bstar < 0 or bstar >= BINS - 1 or not scale > 0:
        return dmax, dmax
    # __fdividef: 2 ulp; model it as the correctly rounded quotient taken two fp32 steps DOWN (the unfavourable side)
    q = F32(bstar + 1) / scale
    q = np.nextafter(np.nextafter(q, F32(0)), F32(0))
    vstar = min(dmax, d0 + float(q * (F32(1) + F32(8) * F32(1.1920929e-7))))
    inb = d[bk == bstar]
    f = ((inb - d0) * float(scale) - float(bstar)) * float(BINS)
    sub = np.where(f < BINS, np.where(f >= 0, f.astype(np.int64), 0), BINS - 1)
    sstar, _ = _kth_bucket(np.bincount(sub, minlength=BINS), k - below)
    assert sstar >= 0                                                   # the bucket holds the k-th smallest by construction
    if sstar >= BINS - 1:
        return vstar, vstar
    v2 = d0 + (float(bstar) + (float(sstar + 1) + 1e-3) / float(BINS)) / float(scale)
    return vstar, min(vstar, v2)


def _check(d, k, d0, dmax):
    kth = np.sort(d)[k - 1] if k <= d.size else np.inf
    vstar, v2 = _select_bounds(d, k, d0, dmax)
    if k > d.size:
        assert vstar == dmax and v2 == dmax                             # fewer than k entries: all of them
        return 0
    assert vstar >= kth, (vstar, kth)
    assert v2 >= kth, (v2, kth)
    # entries the refined bound lets through beyond k (None where the kernel does not refine: last bucket, degenerate range)
    return int((d <= v2).sum() - k) if v2 < vstar else None


@pytest.mark.parametrize("seed", range(8))
def test_bounds_never_fall_below_the_kth_smallest_random(seed):
    rng = np.random.default_rng(seed)
    extra = []
    for _ in range(40):
        n = int(rng.integers(1, 4000))
        k = int(rng.integers(1, 400))
        dmax = float(10.0 ** rng.uniform(-6, 6))
        # distances concentrated below the threshold, as a pruned row's entries are
        d = dmax * (1.0 - rng.random(n) ** float(rng.uniform(1, 8)))
        d0 = 0.0
        if rng.random() < 0.3:                                          # a row without a published threshold: [min, max]
            d0, dmax = float(d.min()), float(d.max())
        extra.append(_check(d, k, d0, dmax))
    # where the second level applies, the bound is the k-th smallest for all practical purposes (2^-20 of the range)
    refined = [e for e in extra if e is not None]
    assert len(refined) > 10 and np.mean(refined) < 1.0


@pytest.mark.parametrize("seed", range(4))
def test_bounds_with_entries_on_bucket_edges_and_duplicates(seed):
    """Adversarial: entries one fp64 / fp32 step either side of coarse and fine bucket edges, heavy duplicates."""
    rng = np.random.default_rng(100 + seed)
    for _ in range(25):
        dmax = float(10.0 ** rng.uniform(-3, 3))
        scale = float(F32(BINS) / F32(dmax))
        edges = rng.integers(1, BINS, size=60) / scale                   # exact coarse edges
        fine = (rng.integers(1, BINS, size=60) + rng.integers(0, BINS, size=60) / BINS) / scale
        pts = np.concatenate([edges, fine])
        near = np.concatenate([pts, np.nextafter(pts, 0.0), np.nextafter(pts, np.inf),
                               np.nextafter(pts.astype(F32), F32(0)).astype(np.float64),
                               np.nextafter(pts.astype(F32), F32(np.inf)).astype(np.float64),
                               np.repeat(pts[:10], 30)])
        d = np.clip(near, 0.0, dmax)
        for k in (1, 7, 100, 250, d.size, d.size + 5):
            _check(d, k, 0.0, dmax)


def test_degenerate_ranges_keep_everything():
    d = np.full(300, 3.25)
    assert _select_bounds(d, 100, 3.25, 3.25) == (3.25, 3.25)           # zero span: one bucket, bound = dmax
    d = 1e9 + np.arange(300) * 1e-3                                      # span below 1e-5 of the values
    v, v2 = _select_bounds(d, 100, float(d.min()), float(d.max()))
    assert v == v2 == float(d.max())
    # NaN / inf distances rank last: they land in the last bucket and never pull the bound down
    d = np.concatenate([np.linspace(0.1, 0.9, 150), [np.inf] * 20, [np.nan] * 20])
    bk = _bucket(d, 0.0, F32(BINS) / F32(1.0))
    assert (bk[150:] == BINS - 1).all() and (bk[:150] < BINS - 1).all()
