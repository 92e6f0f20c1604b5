"""CPU: the error model behind the experimental fp16 tensor-core filter (wc_search_f16.cuh).

The filter may only ever drop a candidate that cannot be among a bin's `refsize` nearest, so its approximate distance
d~ = (n_i + n_j) - 2 * sum_s fp16(x'_i) fp16(x'_j)   (fp32 accumulation, fp32 epilogue)
must stay within the a-priori bound eps * (n_i + n_j) + sub the host code turns into margins, and the margin window around
the k-th smallest approximate distance must contain every true top-k candidate.  Checked here on emulated arithmetic."""
import numpy as np
import pytest

from wisecondor_b200 import synth


def _model(X, chunk=64):
    Xc = X - 1.0
    nrm = (Xc * Xc).sum(axis=1)
    hf = Xc.astype(np.float16).astype(np.float32)
    n32 = nrm.astype(np.float32)
    s = np.zeros((X.shape[0], X.shape[0]), dtype=np.float32)
    for c0 in range(0, X.shape[1], chunk):                       # fp32 accumulation, chunked like the MMA k-steps
        s += hf[:, c0:c0 + chunk] @ hf[:, c0:c0 + chunk].T
    d32 = (n32[:, None] + n32[None, :]) + np.float32(-2.0) * s    # fp32 epilogue
    d_exact = nrm[:, None] + nrm[None, :] - 2.0 * (Xc @ Xc.T)
    return Xc, nrm, d32.astype(np.float64), d_exact


def _eps(S):
    ldh = (S + 63) // 64 * 64
    return 2.0 ** -10 * (1 + 2.0 ** -11) + ldh * 2.0 ** -23 + 2.0 ** -21       # wc_newref_topk: eps16


@pytest.mark.parametrize("S,scale", [(100, 1.0), (600, 1.0), (37, 1e-3), (64, 30.0)])
def test_fp16_filter_error_stays_inside_the_a_priori_bound(S, scale):
    bins = [400, 300, 350]
    X = (synth.corrected_like(bins, S, seed=3) - 1.0) * scale + 1.0
    Xc, nrm, d32, d_exact = _model(X)
    eps = _eps(S)
    sub = 2.0 ** -20 * np.sqrt(S * nrm.max())
    bound = eps * (nrm[:, None] + nrm[None, :]) + sub
    err = np.abs(d32 - d_exact)
    assert (err <= bound).all(), "worst ratio %.3g" % (err / bound).max()
    assert (err / bound).max() < 0.5                               # the bound is meant to be comfortable, not tight


def test_margin_window_contains_the_true_top_k():
    bins = [500, 450, 300, 250]
    S, k = 80, 60
    X = synth.corrected_like(bins, S, seed=8)
    X[700] = X[20]                                                   # exact duplicates across chromosomes
    X[1200] = X[20] + 1e-6
    Xc, nrm, d32, d_exact = _model(X)
    chrom = np.repeat(np.arange(len(bins)), bins)
    eps = _eps(S)
    madd = 2 * eps * nrm.max() + 2.0 ** -20 * np.sqrt(S * nrm.max())
    for i in range(0, X.shape[0], 7):
        other = np.flatnonzero(chrom != chrom[i])
        true_top = other[np.lexsort((other, d_exact[i, other]))[:k]]
        vstar = np.partition(d32[i, other], k - 1)[k - 1]            # k-th smallest APPROXIMATE distance
        window = vstar + 2 * eps * (nrm[i] + abs(vstar)) + madd      # K6's shortlist rule with the fp16 margins
        assert (d32[i, true_top] <= window).all()


def test_values_outside_fp16_are_detected_not_filtered():
    """|x - 1| > 60000 cannot be held in fp16: K4h raises a flag and the call falls back to the fp64 filter."""
    v = np.array([1.0 + 7e4, 1.0 - 7e4], dtype=np.float64) - 1.0
    with np.errstate(over="ignore"):
        assert np.isinf(v.astype(np.float16)).all() and (np.abs(v) > 60000.0).all()
    assert np.isfinite(np.array([6e4], dtype=np.float64).astype(np.float16)).all()
