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
