"""CPU: pins the oracle (oracle/wc_oracle.py, oracle/wc_oracle.c) against golden vectors produced by the reference
itself (tests/golden/make_golden.py ran /root/reference's own code, converted to Python 3, in the build
container).  Bit-exact wherever the reference is deterministic; PCA outputs to 1e-9 relative (LAPACK driver)."""
import os

import numpy as np
import pytest

import c_oracle
import wc_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fn():
    return np.load(os.path.join(GOLD, "functions.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def tiny():
    return np.load(os.path.join(GOLD, "tiny_cli.npz"), allow_pickle=True)


def _close(a, b, rel=1e-9):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    both_nan = np.isnan(a) & np.isnan(b)
    ok = both_nan | (a == b) | (np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)))
    assert ok.all(), "max rel err %g" % np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


# ---- search ------------------------------------------------------------------------------------------------
def test_search_numpy_oracle_bit_exact(fn):
    bins = list(fn['search_bins'])
    X = fn['search_X']
    sums = list(np.cumsum(bins))
    idx, dst = wc_oracle.get_reference(X, bins, sums, 15, 1, 1)
    assert np.array_equal(idx, fn['search_idx'])
    assert np.array_equal(dst, fn['search_dst'])
    pi = [wc_oracle.get_reference(X, bins, sums, 15, p, 4) for p in (1, 2, 3, 4)]
    assert np.array_equal(np.concatenate([p[0] for p in pi]), fn['search_idx'])
    idx, dst = wc_oracle.get_reference(X, bins, sums, 60, 1, 1)
    assert np.array_equal(idx, fn['search_idx_few'])
    assert np.array_equal(dst, fn['search_dst_few'])


def test_search_c_oracle_bit_exact(fn, tiny):
    bins = [int(b) for b in fn['search_bins']]
    idx, dst = c_oracle.get_reference_rows(fn['search_X'], bins, 0, sum(bins), 15)
    assert np.array_equal(idx, fn['search_idx'])
    assert np.array_equal(dst, fn['search_dst'])
    idx, dst = c_oracle.get_reference_rows(fn['search_X'], bins, 0, sum(bins), 60, nthreads=3)
    assert np.array_equal(idx, fn['search_idx_few'])
    assert np.array_equal(dst, fn['search_dst_few'])
    # the reference's CLI run (newrefprep/newrefpart x3/newrefpost) on the tiny genome
    mb = [int(b) for b in tiny['prep_maskedChromBins']]
    idx, dst = c_oracle.get_reference_rows(tiny['prep_correctedData'], mb, 0, sum(mb), 40)
    assert np.array_equal(idx, tiny['ref_indexes'])
    assert np.array_equal(dst, tiny['ref_distances'])


def test_get_part_and_split(fn):
    assert wc_oracle.get_part(0, 3, 100) == (0, 33)
    assert wc_oracle.get_part(2, 3, 100) == (66, 100)
    assert wc_oracle.split_by_chrom(5, 60, [30, 52, 69, 78]) == [[0, 5, 30], [1, 30, 52], [2, 52, 60]]


# ---- ingest / PCA --------------------------------------------------------------------------------------------
def _samples(fn):
    from wisecondor_b200 import synth
    bins = list(fn['ingest_bins'])
    return [synth.counts_to_sample_dict(fn['ingest_counts'][i], bins, 50000000) for i in range(fn['ingest_counts'].shape[0])]


def test_ingest_bit_exact(fn):
    samples = _samples(fn)
    masked, chrom_bins, mask = wc_oracle.to_numpy_array(samples)
    assert np.array_equal(mask, fn['ingest_mask'])
    assert np.array_equal(masked, fn['ingest_masked'])
    scaled = wc_oracle.scale_sample(samples[0], 1, 3)
    assert np.array_equal(scaled['1'], fn['ingest_scaled_chr1']) and scaled['1'].dtype == np.int32
    assert np.array_equal(scaled['2'], fn['ingest_scaled_chr2'])
    sizes = [b + (1 if i % 2 else -1) for i, b in enumerate(chrom_bins)]
    v = wc_oracle.to_numpy_ref_format(samples[1], sizes, np.ones(sum(sizes), dtype=bool))
    assert np.array_equal(v, fn['ingest_padtrunc'])
    assert np.array_equal(wc_oracle.to_numpy_ref_format(samples[2], chrom_bins, mask), fn['ingest_tref'])


def test_pca_matches_full_svd_reference(fn, tiny):
    corrected, comps, mean = wc_oracle.train_pca(fn['ingest_masked'])
    _close(corrected, fn['ingest_corrected'])
    assert np.array_equal(mean, fn['ingest_mean'])        # numpy's pairwise order over each bin's samples
    # components up to a per-row sign (svd_flip convention; does not affect any result)
    for a, b in zip(comps, fn['ingest_components']):
        s = np.sign(np.dot(a, b))
        assert np.max(np.abs(a * s - b)) <= 1e-9 * np.max(np.abs(b))
    _close(wc_oracle.apply_pca(fn['ingest_tref'], fn['ingest_mean'], fn['ingest_components']), fn['ingest_applied'], 1e-12)
    corrected, comps, mean = wc_oracle.train_pca(tiny['prep_maskedData'])
    _close(corrected, tiny['prep_correctedData'])
    assert np.array_equal(mean, tiny['ref_pca_mean'])


# ---- z-scores ----------------------------------------------------------------------------------------------
def test_zscores_bit_exact(fn):
    bins = list(fn['z_bins'])
    sums = list(np.cumsum(bins))
    cutoff = wc_oracle.get_optimal_cutoff(fn['z_dst'], 3)
    assert cutoff == float(fn['z_cutoff'])
    test = fn['z_test']
    z, r, s, sd = wc_oracle.try_sample(test, np.copy(test), fn['z_idx'], fn['z_dst'], bins, sums, cutoff)
    assert np.array_equal(np.array([z, r, s]), fn['z_pass1'], equal_nan=True) and sd == float(fn['z_pass1_sd'])
    z, r, s, sd = wc_oracle.repeat_test(np.copy(test), fn['z_idx'], fn['z_dst'], bins, sums, cutoff, 3.0, 5)
    assert np.array_equal(np.array([z, r, s]), fn['z_pass5'], equal_nan=True) and sd == float(fn['z_pass5_sd'])


def test_pairwise_model_is_numpy(fn):
    """The summation-order model the CUDA z-score kernel implements (wc_oracle.pairwise_sum) is numpy's."""
    rng = np.random.default_rng(3)
    for n in list(range(0, 40)) + [64, 100, 127, 128, 129, 200, 255, 256, 257, 384, 1000]:
        a = rng.normal(1.0, 0.1, size=n)
        assert wc_oracle.pairwise_sum(a) == (np.sum(a) if n else 0.0)
        if n:
            m, sd = wc_oracle.mean_std_model(a)
            assert m == np.mean(a) and sd == np.std(a)


# ---- segmentation --------------------------------------------------------------------------------------------
def test_segmentation_bit_exact(fn):
    offs = np.concatenate(([0], np.cumsum(fn['seg_n'])))
    calls = fn['seg_calls']
    for i in range(len(fn['seg_n'])):
        zreg = fn['seg_z'][offs[i]:offs[i + 1]]
        cw, segs = wc_oracle.segment_region(zreg, 3.5, 3)
        assert cw == fn['seg_cw'][i]
        want = calls[calls[:, 0] == i][:, 1:]
        got = np.array([[s[1][0], s[1][1], s[0]] for s in segs], dtype=float).reshape(-1, 3)
        assert np.array_equal(got, want), "region %d" % i
        # the relaxed prefix-sum oracle picks the same runs
        cw2, segs2 = wc_oracle.segment_region_prefix(zreg, 3.5, 3)
        assert [s[1] for s in segs2] == [s[1] for s in segs]
        _close([s[0] for s in segs2] + [cw2], [s[0] for s in segs] + [cw], 1e-12)
    cw, segs = wc_oracle.segment_region(np.full(25, 2.0), 3.5, 3)
    assert np.array_equal(np.array([[s[1][0], s[1][1], s[0]] for s in segs], dtype=float), fn['seg_flat_calls'])


def test_segmentation_min_effect(fn):
    cw, segs = wc_oracle.segment_region(fn['segmin_z'], 3.5, 3, fn['segmin_r'], 0.05)
    assert cw == float(fn['segmin_cw'])
    assert np.array_equal(np.array([[s[1][0], s[1][1], s[0]] for s in segs], dtype=float).reshape(-1, 3),
                          fn['segmin_calls'].reshape(-1, 3))


def _segmin_nonfinite():
    import sys
    import warnings
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "segmin_nonfinite.npz"))
    return [(name, z, r, gold[name + "_cw"], gold[name + "_calls"]) for name, z, r in make_golden.segmin_nonfinite_cases()]


def test_segmentation_min_effect_non_finite():
    """inf / NaN z-scores and ratios under the effect-size filter: the reference lets them flow through fillTriMin and
    segmentTri under np.seterr('ignore') (wisetools.py:475-487, triarray.py:59-84); golden = its own output."""
    for name, z, r, cw_want, calls_want in _segmin_nonfinite():
        with np.errstate(all="ignore"):
            cw, segs = wc_oracle.segment_region(z, 3.5, 3, r, 0.05)
        assert np.array_equal(np.array([cw]), np.array([cw_want]), equal_nan=True), name
        got = np.array([[s[1][0], s[1][1], s[0]] for s in segs], dtype=float).reshape(-1, 3)
        assert np.array_equal(got, calls_want.reshape(-1, 3), equal_nan=True), name


# ---- the whole test tool ---------------------------------------------------------------------------------------
def test_tool_test_matches_reference_cli(tiny):
    from wisecondor_b200 import synth
    bins = [int(b) for b in tiny['bins']]
    ref = dict(binsize=tiny['ref_binsize'].item(), indexes=tiny['ref_indexes'], distances=tiny['ref_distances'],
               chromosome_sizes=tiny['ref_chromosome_sizes'], mask=tiny['ref_mask'], masked_sizes=tiny['ref_masked_sizes'],
               pca_mean=tiny['ref_pca_mean'], pca_components=tiny['ref_pca_components'])
    for t in range(tiny['test_counts'].shape[0]):
        sample = synth.counts_to_sample_dict(tiny['test_counts'][t], bins, 50000000)
        res = wc_oracle.test_sample(sample, float(tiny['binsize']), ref, minrefbins=10, repeats=3 if t == 2 else 5)
        assert np.array_equal(np.concatenate(res['results_z']), tiny['res%d_z' % t], equal_nan=True)
        assert np.array_equal(np.concatenate(res['results_r']), tiny['res%d_r' % t], equal_nan=True)
        assert np.array_equal(res['results_cwz'], tiny['res%d_cwz' % t])
        assert np.array_equal(np.asarray(res['results_calls'], dtype=float).reshape(-1, 5), tiny['res%d_calls' % t])
        assert np.array_equal(np.array([res['threshold_z'], res['asdef'], res['aasdef']]), tiny['res%d_scalars' % t])
