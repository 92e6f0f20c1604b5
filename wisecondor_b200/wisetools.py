"""Host-side mirror of the reference's numerics library (/root/reference/wisetools.py) for the hot path.

Same function names, argument meaning and return values as the reference; the bodies call the CUDA library
through wisecondor_b200._cabi / wisecondor_b200.device.  Functions the north star keeps on the host
(scaleSample, getOptimalCutoff, getPart, splitByChrom, inflateArray) are plain numpy.
"""
import numpy as np

from . import device as _dev

DEVICE = 0   # CUDA device the host-facing functions use


def getPart(partnum, outof, bincount):
    """Start and end bin of 0-based part `partnum` of `outof` (reference wisetools.py:358-361)."""
    return int(bincount / float(outof) * partnum), int(bincount / float(outof) * (partnum + 1))


def getReference(correctedData, chromosomeBins, chromosomeBinSums, selectRefAmount=100, part=1, splitParts=1):
    """Reference bins for the rows of 1-based `part` of `splitParts` (reference wisetools.py:364-398).

    Returns (indexArray int32 rows x selectRefAmount, distanceArray float64 rows x selectRefAmount): per target
    bin the positions, within the concatenation of all other chromosomes, of the selectRefAmount nearest bins
    ordered by (squared distance, index), and those distances.  The chromosome split of the reference's
    splitByChrom/getRefForBins loop happens inside the kernel (per-row exclusion ranges).
    """
    bincount = int(chromosomeBinSums[-1])
    startNum, endNum = getPart(part - 1, splitParts, bincount)
    print('Working on part', part, 'of', splitParts, 'meaning bins', startNum, 'up to', endNum)
    X = np.ascontiguousarray(correctedData, dtype=np.float64)
    if X.shape[0] != bincount:
        raise ValueError("correctedData has %d bins, chromosomeBinSums says %d" % (X.shape[0], bincount))
    idx, dist = _dev.newref_topk_host(X, [int(b) for b in chromosomeBins], startNum, endNum,
                                      int(selectRefAmount), device=DEVICE)
    return idx, dist


def getOptimalCutoff(reference, repeats):
    """mean + 3 sigma of the distances below the previous cutoff, `repeats` rounds (reference
    wisetools.py:328-336).  Sample independent, computed once per reference on the host with numpy so the
    value is bit-identical to the reference's."""
    optimalCutoff = float("inf")
    mask = np.zeros(reference.shape)
    for _ in range(repeats):
        mask = reference < optimalCutoff
        optimalCutoff = np.average(reference[mask]) + 3 * np.std(reference[mask])
    return optimalCutoff, mask


def scaleSample(sample, fromSize, toSize):
    """Down-bin a sample dict by an integer factor (reference wisetools.py:220-237).  Host."""
    if fromSize == toSize or toSize is None:
        return sample
    if toSize == 0 or fromSize == 0 or toSize < fromSize or toSize % fromSize > 0:
        print('ERROR: Impossible binsize scaling requested:', fromSize, 'to', toSize)
        raise SystemExit(1)
    scale = int(toSize // fromSize)
    out = dict()
    for chrom in sample:
        data = np.asarray(sample[chrom])
        newLen = int(np.ceil(len(data) / float(scale)))
        padded = np.zeros(newLen * scale, dtype=np.int64)
        padded[:len(data)] = data
        out[chrom] = padded.reshape(newLen, scale).sum(axis=1).astype(np.int32)
    return out


def inflateArray(array, mask):
    """Scatter `array` into the True positions of `mask`, zeros elsewhere (reference wisetools.py:281-288)."""
    temp = np.zeros(mask.shape[0])
    temp[np.asarray(mask, dtype=bool)] = array
    return temp


def inflateArrayMulti(array, mask_list):
    """reference wisetools.py:291-295."""
    temp = array
    for mask in reversed(mask_list):
        temp = inflateArray(temp, mask)
    return temp


# ------------------------------------------------------------------------------------------------------------
# test path
# ------------------------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def _refFormatCounts(sample, chromBins):
    """Host part of toNumpyRefFormat (reference wisetools.py:268-273): per autosome zero-pad / truncate the
    count array to chromBins[c] and concatenate.  Returns int32 [Nraw]."""
    parts = []
    for chromosome in range(1, 23):
        want = int(chromBins[chromosome - 1])
        have = np.asarray(sample[str(chromosome)])
        thisChrom = np.zeros(want, dtype=np.int32)
        minLen = min(want, len(have))
        thisChrom[:minLen] = have[:minLen]
        parts.append(thisChrom)
    return np.concatenate(parts)


def toNumpyRefFormat(sample, chromBins, mask):
    """reference wisetools.py:267-278: masked, total-normalised bin vector of one sample (numpy float64 [N])."""
    return prepSamples([sample], chromBins, mask, None, None)[:, 0].copy()


def applyPCA(sampleData, mean, components):
    """reference wisetools.py:104-113: sampleData / ((sampleData - mean) C^T C + mean)."""
    torch = _torch()
    dev = torch.device("cuda", DEVICE)
    x = torch.as_tensor(np.ascontiguousarray(sampleData, dtype=np.float64)[None, :], device=dev)
    out = _dev.apply_pca(x, torch.as_tensor(np.ascontiguousarray(mean, dtype=np.float64), device=dev),
                         torch.as_tensor(np.ascontiguousarray(components, dtype=np.float64), device=dev))
    return out[:, 0].cpu().numpy()


def prepSample(sample, chromosome_sizes, mask, pca_mean, pca_components):
    """reference wisetools.py:401-404."""
    return prepSamples([sample], chromosome_sizes, mask, pca_mean, pca_components)[:, 0].copy()


def prepSamples(samples, chromosome_sizes, mask, pca_mean, pca_components, as_device=False):
    """prepSample for a batch: returns T [N][B'] (bin-major, sample-minor; B' = B rounded up to 32 when
    as_device, else exactly B as numpy)."""
    torch = _torch()
    dev = torch.device("cuda", DEVICE)
    counts = np.stack([_refFormatCounts(s, chromosome_sizes) for s in samples])
    masked_raw = np.flatnonzero(np.asarray(mask, dtype=bool)).astype(np.int32)
    pm = pc = None
    if pca_components is not None:
        pm = torch.as_tensor(np.ascontiguousarray(pca_mean, dtype=np.float64), device=dev)
        pc = torch.as_tensor(np.ascontiguousarray(pca_components, dtype=np.float64), device=dev)
    T = _dev.test_prep(torch.as_tensor(counts, device=dev), torch.as_tensor(masked_raw, device=dev), pm, pc)
    if as_device:
        return T
    return T[:, :len(samples)].cpu().numpy()


_TABLE_CACHE = {}


def _table(indexes, distances, chromosomeBins, cutoff):
    key = (id(indexes), id(distances), float(cutoff), DEVICE)
    hit = _TABLE_CACHE.get(key)
    if hit is None or hit[0] is not indexes:
        _TABLE_CACHE.clear()
        hit = (indexes, _dev.ReferenceTable(indexes, distances, chromosomeBins, cutoff, device=DEVICE))
        _TABLE_CACHE[key] = hit
    return hit[1]


def repeatTestBatch(testData, indexes, distances, chromosomeBins, chromosomeBinSums, cutoff, threshold, repeats):
    """repeatTest for many samples at once.  testData: numpy [B][N] (one corrected sample per row) or a CUDA
    tensor [N][ldb] from prepSamples(as_device=True) together with B = testData.nsamples... plain numpy here.
    Returns (resultsZ [B][N], resultsR [B][N], refSizes [B][N] float, stdDevAvg [B])."""
    torch = _torch()
    dev = torch.device("cuda", DEVICE)
    X = np.ascontiguousarray(testData, dtype=np.float64)
    b, n = X.shape
    ldb = _dev.pad32(b)
    T = torch.ones((n, ldb), dtype=torch.float64, device=dev)
    T[:, :b] = torch.as_tensor(X, device=dev).T
    table = _table(indexes, distances, chromosomeBins, cutoff)
    z, r, sizes, asdef = _dev.zscore_batch(T, b, table, threshold, repeats)
    return z.cpu().numpy(), r.cpu().numpy(), sizes.cpu().numpy().astype(float), asdef.cpu().numpy()


def trySample(testData, testCopy, indexes, distances, chromosomeBins, chromosomeBinSums, cutoff):
    """reference wisetools.py:407-435: one z-score pass.  `testCopy` carries the -1 marks of earlier passes."""
    torch = _torch()
    dev = torch.device("cuda", DEVICE)
    n = len(testData)
    T = torch.ones((n, 32), dtype=torch.float64, device=dev)
    C = torch.ones((n, 32), dtype=torch.float64, device=dev)
    T[:, 0] = torch.as_tensor(np.ascontiguousarray(testData, dtype=np.float64), device=dev)
    C[:, 0] = torch.as_tensor(np.ascontiguousarray(testCopy, dtype=np.float64), device=dev)
    table = _table(indexes, distances, chromosomeBins, cutoff)
    z, r, sizes, asdef = _dev.zscore_batch(T, 1, table, float("inf"), 1, copy_init=C)
    return z[0].cpu().numpy(), r[0].cpu().numpy(), sizes[0].cpu().numpy().astype(float), float(asdef[0].item())


def repeatTest(testData, indexes, distances, chromosomeBins, chromosomeBinSums, cutoff, threshold, repeats):
    """reference wisetools.py:438-448 for one sample."""
    z, r, sizes, asdef = repeatTestBatch(np.asarray(testData)[None, :], indexes, distances, chromosomeBins,
                                         chromosomeBinSums, cutoff, threshold, repeats)
    return z[0], r[0], sizes[0], float(asdef[0])


def segmentChromosomes(cleanedZ_or_z, refSizes, masked_sizes, chromosomes, minrefbins, z_threshold, min_search=3):
    """fillTri + segmentTri for the listed chromosomes (1-based, as -chromosomes) of a batch (reference
    wisecondor.py:233-238, wisetools.py:466-472, triarray.py:59-84).  z, refSizes: numpy [B][N].
    Returns (chromWide [B][nsel], cleanedBins [B][nsel], calls structured array sorted by (sample, chrom, x))."""
    torch = _torch()
    dev = torch.device("cuda", DEVICE)
    z = torch.as_tensor(np.ascontiguousarray(cleanedZ_or_z, dtype=np.float64), device=dev)
    sizes = torch.as_tensor(np.ascontiguousarray(refSizes).astype(np.int32), device=dev)
    cwz, cleaned, calls = _dev.segment_batch(z, sizes, masked_sizes, [c - 1 for c in chromosomes], minrefbins,
                                             z_threshold, min_search)
    return cwz.cpu().numpy(), cleaned.cpu().numpy(), calls
