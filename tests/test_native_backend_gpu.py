"""GPU: the PyTorch-free buffer backend of the host layer (wisecondor_b200._mem.DevArray over wc_dev_alloc / wc_copy_*).
It must give the results of the torch-tensor backend bit for bit (same kernels, same pointers), and the command line
must be able to run a whole `newref` and `test` without ever importing torch."""
import os
import subprocess
import sys

import numpy as np
import pytest

import wc_oracle
from wisecondor_b200 import _mem, device, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_devarray_round_trip_and_checks():
    rng = np.random.default_rng(0)
    for shape, dtype in (((0,), np.float64), ((7, 33), np.float64), ((5, 3), np.int32), ((100,), np.uint8)):
        a = (rng.normal(size=shape) * 100).astype(dtype)
        d = _mem.DevArray.from_numpy(a)
        assert d.shape == a.shape and d.dtype == a.dtype and d.is_contiguous()
        assert np.array_equal(d.numpy(), a)
    with pytest.raises(Exception, match="device memory"):
        device.newref_topk(np.zeros((4, 4)), [2, 2], 0, 4, 1)
    with pytest.raises(Exception, match="contiguous"):
        device.newref_topk(_mem.DevArray.from_numpy(np.zeros((4, 4), dtype=np.int32)), [2, 2], 0, 4, 1)


def test_search_native_equals_torch_and_oracle():
    import torch
    bins = [40, 31, 50, 23]
    X = synth.corrected_like(bins, 24, seed=11)
    want_idx, want_dst = wc_oracle.get_reference(X, bins, list(np.cumsum(bins)), 20, 1, 1)
    ni, nd = device.newref_topk(_mem.DevArray.from_numpy(X), bins, 0, sum(bins), 20)
    assert isinstance(ni, _mem.DevArray) and isinstance(nd, _mem.DevArray)
    ti, td = device.newref_topk(torch.as_tensor(X, device="cuda:0"), bins, 0, sum(bins), 20)
    assert np.array_equal(ni.numpy(), want_idx) and np.array_equal(nd.numpy(), want_dst)
    assert np.array_equal(ti.cpu().numpy(), want_idx) and np.array_equal(td.cpu().numpy(), want_dst)


def test_whole_test_path_native_equals_torch(monkeypatch):
    from wisecondor_b200 import wisetools
    tiny = np.load(os.path.join(GOLD, "tiny_cli.npz"), allow_pickle=True)
    bins = [int(b) for b in tiny['bins']]
    ref = dict(binsize=tiny['ref_binsize'].item(), indexes=tiny['ref_indexes'], distances=tiny['ref_distances'],
               chromosome_sizes=tiny['ref_chromosome_sizes'], mask=tiny['ref_mask'], masked_sizes=tiny['ref_masked_sizes'],
               pca_mean=tiny['ref_pca_mean'], pca_components=tiny['ref_pca_components'])
    samples = [synth.counts_to_sample_dict(tiny['test_counts'][t], bins, 50000000) for t in range(tiny['test_counts'].shape[0])]
    got = {}
    for backend in ("torch", "native"):
        monkeypatch.setattr(_mem, "BACKEND", backend)
        wisetools._TABLE_CACHE.clear()
        got[backend] = wisetools.testSamples(samples, ref, float(tiny['res0_scalars'][0]), minrefbins=10, batch=3)
    for a, b in zip(got["torch"], got["native"]):
        for key in ("results_z", "results_r"):
            assert np.array_equal(np.concatenate(a[key]), np.concatenate(b[key]), equal_nan=True)
        assert np.array_equal(a["results_cwz"], b["results_cwz"])
        assert np.array_equal(np.asarray(a["results_calls"]), np.asarray(b["results_calls"]))
        assert a["asdef"] == b["asdef"]
    # and both are the reference's numbers
    for t, res in enumerate(got["native"]):
        if t == 2:
            continue                                   # golden case 2 was produced with -repeats 3
        assert np.allclose(np.concatenate(res['results_z']), tiny['res%d_z' % t], rtol=1e-9, atol=0, equal_nan=True)
        calls = np.asarray(res['results_calls'], dtype=float).reshape(-1, 5)
        assert np.array_equal(calls[:, :3], tiny['res%d_calls' % t][:, :3])


def test_prep_native_equals_torch(monkeypatch):
    from wisecondor_b200 import wisetools
    fn = np.load(os.path.join(GOLD, "functions.npz"), allow_pickle=True)
    bins = list(fn['ingest_bins'])
    samples = [synth.counts_to_sample_dict(fn['ingest_counts'][i], bins, 50000000) for i in range(fn['ingest_counts'].shape[0])]
    out = {}
    for backend in ("torch", "native"):
        monkeypatch.setattr(_mem, "BACKEND", backend)
        masked, chrom_bins, mask = wisetools.toNumpyArray(samples, as_device=True)
        corrected, pca = wisetools.trainPCA(masked, as_device=True)
        assert _mem.is_native(corrected) == (backend == "native")
        out[backend] = (_mem.to_host(masked), mask, _mem.to_host(corrected), pca.components_, pca.mean_)
    for a, b in zip(out["torch"], out["native"]):
        assert np.array_equal(a, b)
    assert np.array_equal(out["native"][0], fn['ingest_masked'])


def test_command_line_runs_without_torch(tmp_path):
    """newref (prep, parts, post) and test in fresh interpreters: results equal the reference's golden run and `torch`
    never enters sys.modules."""
    sys.path.insert(0, GOLD)
    from make_golden import TINY_BINSIZE, write_sample_npz
    tiny = np.load(os.path.join(GOLD, "tiny_cli.npz"), allow_pickle=True)
    bins = [int(b) for b in tiny['bins']]
    refs = []
    for i in range(tiny['ref_counts'].shape[0]):
        refs.append(str(tmp_path / ("r%02d.npz" % i)))
        write_sample_npz(refs[-1], tiny['ref_counts'][i], bins, TINY_BINSIZE)
    write_sample_npz(str(tmp_path / "t0.npz"), tiny['test_counts'][0], bins, TINY_BINSIZE)
    code = ("import sys; sys.path.insert(0, %r); import wisecondor; wisecondor.main(sys.argv[1:]); "
            "assert 'torch' not in sys.modules, 'PyTorch was imported'") % ROOT
    env = dict(os.environ)
    env.pop("WISECONDOR_BACKEND", None)

    def run(*argv):
        r = subprocess.run([sys.executable, "-c", code] + list(argv), capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]

    run("newref", *(refs + [str(tmp_path / "ref.npz"), "-refsize", "40", "-parts", "3"]))
    ref = np.load(str(tmp_path / "ref.npz"), allow_pickle=True)
    assert np.array_equal(ref['indexes'], tiny['ref_indexes'])
    assert np.allclose(ref['distances'], tiny['ref_distances'], rtol=1e-9, atol=0)    # PCA by eigh here, full SVD there
    assert np.array_equal(ref['pca_mean'], tiny['ref_pca_mean'])
    run("test", str(tmp_path / "t0.npz"), str(tmp_path / "o0.npz"), str(tmp_path / "ref.npz"), "-minrefbins", "10")
    res = np.load(str(tmp_path / "o0.npz"), allow_pickle=True)
    assert np.allclose(np.concatenate(list(res['results_z'])), tiny['res0_z'], rtol=1e-6, atol=1e-9, equal_nan=True)
