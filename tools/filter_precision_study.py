"""CPU study of reduced-precision tensor-core filters in front of the exact fp64 re-score (K5 successors): a bf16 x 3
split and the single-pass fp16 scheme K5h implements (wc_search_f16.cuh).

x' = x - 1 is split into two bf16 numbers h + l (|x' - h - l| <= 2^-17 |x'|); the filter score is
    s~ = sum_s (h_i h_j + h_i l_j + l_i h_j)      accumulated in fp32 in chunks of `chunk` samples, chunks added in fp64,
and d~ = n_i + n_j - 2 s~.  Prints the observed error of d~ against the exact fp64 distance, relative to (n_i + n_j),
and how many candidates per row a margin m * (n_i + n_j) around the k-th smallest distance lets through."""
import sys
import numpy as np

sys.path.insert(0, ".")
from wisecondor_b200 import synth


def bf16(a):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).view(np.float32)


def main():
    binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    rows = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    bins = synth.chrom_bins(binsize)
    X = synth.corrected_like(bins, S, seed=4)
    n = X.shape[0]
    Xc = X - 1.0
    nrm = (Xc * Xc).sum(axis=1)
    h = bf16(Xc)
    l = bf16(Xc - h.astype(np.float64))
    split_err = np.abs(Xc - h.astype(np.float64) - l.astype(np.float64)).max() / np.abs(Xc).max()
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(n, size=rows, replace=False))
    chrom = np.repeat(np.arange(len(bins)), bins)
    # exact
    d_exact = nrm[sel][:, None] + nrm[None, :] - 2.0 * (Xc[sel] @ Xc.T)
    # bf16x3 with fp32 accumulation inside chunks (numpy float32 matmul: pairwise/blocked fp32 - an optimistic model of
    # the tensor core's accumulator; the bound below does not rely on it) and fp64 across chunks
    s = np.zeros((rows, n))
    for c0 in range(0, S, chunk):
        hs, ls = h[sel, c0:c0 + chunk], l[sel, c0:c0 + chunk]
        ha, la = h[:, c0:c0 + chunk], l[:, c0:c0 + chunk]
        s += (hs @ ha.T + hs @ la.T + ls @ ha.T).astype(np.float64)
    d_apx = nrm[sel][:, None] + nrm[None, :] - 2.0 * s
    scale = nrm[sel][:, None] + nrm[None, :]
    rel = np.abs(d_apx - d_exact) / scale
    print("bins %d S %d k %d rows %d chunk %d" % (n, S, k, rows, chunk))
    print("split error |x'-h-l|/max|x'|: %.3g (2^-17 = %.3g)" % (split_err, 2.0 ** -17))
    print("filter error |d~ - d| / (n_i + n_j): max %.3g  mean %.3g" % (rel.max(), rel.mean()))
    # worst-case bound: dropped l*l and split remainders 3 * 2^-17 * |x_i||x_j| + fp32 accumulation of 3*chunk products
    bound = (3 * 2.0 ** -17 + 3 * chunk * 2.0 ** -24 + (S / chunk) * 2.0 ** -53) * 0.5 * 2.0
    print("a priori bound on it (split + %d-term fp32 accumulation): %.3g" % (3 * chunk, bound))
    other = chrom[sel][:, None] != chrom[None, :]
    for m in (0.0, 1e-5, bound, 1e-4, 3e-4, 1e-3, 4e-3):
        counts = []
        for r in range(rows):
            d = d_exact[r][other[r]]
            kth = np.partition(d, k - 1)[k - 1]
            counts.append(int((d <= kth + 2 * m * (nrm[sel[r]] + np.max(nrm))).sum()))     # margin on both sides of the filter
        counts = np.array(counts)
        print("margin %.2g * (n_i + n_max): candidates per row mean %.1f  max %d  (k = %d)" % (m, counts.mean(), counts.max(), k))
    kth_all = np.array([np.partition(d_exact[r][other[r]], k - 1)[k - 1] for r in range(rows)])
    # the scheme K5h implements: ONE fp16 rounding of x' (fp32 accumulation), margin 2 * eps * (n_i + n_max)
    hf = Xc.astype(np.float16).astype(np.float32)
    s1 = np.zeros((rows, n))
    for c0 in range(0, S, chunk):
        s1 += (hf[sel, c0:c0 + chunk] @ hf[:, c0:c0 + chunk].T).astype(np.float64)
    d1 = nrm[sel][:, None] + nrm[None, :] - 2.0 * s1
    rel1 = np.abs(d1 - d_exact) / scale
    eps16 = 2.0 ** -10 * (1 + 2.0 ** -11) + ((S + 63) // 64 * 64) * 2.0 ** -23 + 2.0 ** -21
    print("single-pass fp16: error / (n_i + n_j): max %.3g mean %.3g ; a priori eps %.3g" % (rel1.max(), rel1.mean(), eps16))
    cnt = []
    for r in range(rows):
        d = d1[r][other[r]]
        kth = np.partition(d, k - 1)[k - 1]
        cnt.append(int((d <= kth + 2 * eps16 * (nrm[sel[r]] + abs(kth)) + 2 * eps16 * np.max(nrm)).sum()))
    cnt = np.array(cnt)
    print("single-pass fp16: shortlist per row with K6's window (on the approximate distances): mean %.1f max %d (k = %d)"
          % (cnt.mean(), cnt.max(), k))
    print("k-th distance / n_i: median %.3g ; n_i median %.3g" % (np.median(kth_all / nrm[sel]), np.median(nrm)))


if __name__ == "__main__":
    main()
