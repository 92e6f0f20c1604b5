"""Synthetic WISECONDOR inputs: seeded sample x bin read-count matrices.

Bins per chromosome follow the reference's convert step, `int(length / binsize + 1)`
(/root/reference/wisetools.py:152), over the hg19 chromosome lengths.  Counts are
Poisson(lambda_b * exp(w_s . f_b)): a per-bin base rate shared by all samples, three per-bin latent
factors and per-sample weights, so that the 3-component PCA of newref has something real to remove.
Reference and test samples must be drawn with the same `bin_seed` (same lambda_b, f_b).
"""
import numpy as np

HG19_LENGTHS = {
    '1': 249250621, '2': 243199373, '3': 198022430, '4': 191154276, '5': 180915260, '6': 171115067,
    '7': 159138663, '8': 146364022, '9': 141213431, '10': 135534747, '11': 135006516, '12': 133851895,
    '13': 115169878, '14': 107349540, '15': 102531392, '16': 90354753, '17': 81195210, '18': 78077248,
    '19': 59128983, '20': 63025520, '21': 48129895, '22': 51304566, 'X': 155270560, 'Y': 59373566,
}
AUTOSOMES = [str(c) for c in range(1, 23)]


def chrom_bins(binsize, chroms=AUTOSOMES, lengths=HG19_LENGTHS):
    """Bins per chromosome, as convert allocates them (wisetools.py:152)."""
    return [int(lengths[c] / float(binsize) + 1) for c in chroms]


def bin_model(binsize, bin_seed=1, depth_per_mb=3480.0, zero_frac=0.05, scale_bins=None):
    """Per-bin base rates and latent factors shared by every sample of one 'laboratory'.

    `scale_bins` (list of 22 ints) overrides the hg19-derived autosome bin counts: tests use it to build small
    genomes with the same structure.  Returns (bins[22], lam[Nraw], fac[Nraw,3]).
    """
    rng = np.random.default_rng(bin_seed)
    bins = list(scale_bins) if scale_bins is not None else chrom_bins(binsize)
    nraw = int(sum(bins))
    depth = depth_per_mb * binsize / 1e6
    lam = rng.gamma(20.0, depth / 20.0, size=nraw)
    fac = rng.normal(0.0, 0.05, size=(nraw, 3))
    if zero_frac > 0:
        nblocks = max(1, int(nraw * zero_frac / 8))
        starts = rng.integers(0, nraw, size=nblocks)
        for s in starts:
            lam[s:s + int(rng.integers(1, 16))] = 0.0
    return bins, lam, fac


def sample_counts(nsamples, lam, fac, seed=2, dtype=np.int32):
    """Count matrix [nsamples][Nraw] (sample-major, the layout of per-sample npz arrays)."""
    rng = np.random.default_rng(seed)
    w = rng.normal(0.0, 1.0, size=(nsamples, 3))
    out = np.empty((nsamples, lam.shape[0]), dtype=dtype)
    for s in range(nsamples):
        rate = lam * np.exp(fac @ w[s])
        out[s] = rng.poisson(rate)
    return out


def counts_to_sample_dict(row, bins, binsize, xy_seed=None):
    """One row of `sample_counts` -> the `sample` dict convert writes ('1'..'22','X','Y' -> int32 arrays)."""
    sample = {}
    pos = 0
    for c, n in zip(AUTOSOMES, bins):
        sample[c] = np.ascontiguousarray(row[pos:pos + n]).astype(np.int32)
        pos += n
    for c in ('X', 'Y'):
        n = int(HG19_LENGTHS[c] / float(binsize) + 1)
        rng = np.random.default_rng(0 if xy_seed is None else xy_seed)
        sample[c] = rng.poisson(10.0, size=n).astype(np.int32)
    return sample


def inject_aberration(row, bins, chrom, start_frac, end_frac, factor, seed=3):
    """Scale the counts of a chromosome stretch by `factor` (binomial thinning / Poisson boosting)."""
    rng = np.random.default_rng(seed)
    off = int(sum(bins[:chrom - 1]))
    a = off + int(bins[chrom - 1] * start_frac)
    b = off + max(int(bins[chrom - 1] * end_frac), int(bins[chrom - 1] * start_frac) + 1)
    seg = row[a:b].astype(np.int64)
    if factor <= 1.0:
        row[a:b] = rng.binomial(seg, factor)
    else:
        row[a:b] = seg + rng.poisson(seg * (factor - 1.0))
    return a - off, b - off


def corrected_like(nbins_per_chrom, nsamples, seed=5, sigma=0.05, dtype=np.float64):
    """A PCA-corrected-looking matrix (values ~1) generated directly: [N][S], C-order.  Used for the search
    kernels' large-shape tests and the bench, where running normalise+PCA first is not the point."""
    rng = np.random.default_rng(seed)
    n = int(sum(nbins_per_chrom))
    # a few bin 'families' so that nearest neighbours are meaningful, plus independent noise
    nfam = max(8, n // 64)
    fam = rng.integers(0, nfam, size=n)
    fam_profile = rng.normal(0.0, sigma, size=(nfam, nsamples))
    x = 1.0 + fam_profile[fam] * rng.uniform(0.3, 1.0, size=(n, 1)) + rng.normal(0.0, sigma, size=(n, nsamples))
    return np.ascontiguousarray(x.astype(dtype))


def corrected_like_device(nbins_per_chrom, nsamples, seed=5, sigma=0.05, device=None, chunk=8192):
    """Device-side counterpart of corrected_like for matrices too large to build on the host (2000 x 10 kb = 4.6 GB):
    same structure (bin families + independent noise), generated in row chunks with a seeded torch generator, so
    every rank of one job - same GPU model, same seed - holds the same matrix.  Not bit-equal to corrected_like."""
    import torch
    n = int(sum(nbins_per_chrom))
    dev = torch.device("cuda") if device is None else device
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    nfam = max(8, n // 64)
    fam_profile = torch.randn((nfam, nsamples), generator=g, device=dev, dtype=torch.float64) * sigma
    out = torch.empty((n, nsamples), dtype=torch.float64, device=dev)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        fam = torch.randint(0, nfam, (b - a,), generator=g, device=dev)
        scale = torch.rand((b - a, 1), generator=g, device=dev, dtype=torch.float64) * 0.7 + 0.3
        noise = torch.randn((b - a, nsamples), generator=g, device=dev, dtype=torch.float64) * sigma
        out[a:b] = 1.0 + fam_profile[fam] * scale + noise
    return out
