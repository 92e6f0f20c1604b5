"""CPU: host-side logic (row/sample partition, scaling, inflate, CLI surface), the C-ABI library (loads, exports every
symbol the header declares, fails loudly without a GPU) and the world_size-2 gather over gloo."""
import ctypes
import os
import re

import numpy as np
import pytest

import wc_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from wisecondor_b200 import _cabi, build
    path = build.build_library()
    L = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "wisecondor_b200.h")).read()
    declared = set(re.findall(r"\b(wc_[a-z0-9_]+)\s*\(", header))
    declared -= {"wc_call", "wc_ctx", "wc_status"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), "libwisecondor_b200.so lacks %s" % name
    assert set(_cabi.SYMBOLS) == declared, "binding list and header differ: %s" % (set(_cabi.SYMBOLS) ^ declared)
    assert b"sm_100a" in _cabi.lib().wc_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wisecondor_b200 import _cabi, wisetools
    with pytest.raises(_cabi.WisecondorError):
        _cabi.Context(0)
    X = np.ones((10, 4))
    with pytest.raises(Exception):
        wisetools.getReference(X, [5, 5], [5, 10], 3, 1, 1)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no product source imports, loads or links it."""
    bad = re.compile(r"^\s*(import|from)\s+\S*(wc_oracle|c_oracle)|libwc_oracle|sys\.path.*oracle|#include.*oracle", re.M)
    files = [os.path.join(ROOT, "wisecondor.py")]
    for dirpath, _, names in os.walk(os.path.join(ROOT, "wisecondor_b200")):
        files += [os.path.join(dirpath, f) for f in names if f.endswith((".py", ".cu", ".cuh", ".h"))]
    for f in files:
        assert not bad.search(open(f).read()), f


def test_get_part_matches_reference_formula():
    from wisecondor_b200 import shard, wisetools
    for n in (7, 716, 11537, 57633, 288113):
        for parts in (1, 2, 3, 4, 7, 8):
            edges = [wisetools.getPart(p, parts, n) for p in range(parts)]
            assert edges == [wc_oracle.get_part(p, parts, n) for p in range(parts)]
            assert edges == [shard.row_shard(p, parts, n) for p in range(parts)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(parts - 1))
    for n in (0, 1, 5, 10000):
        for w in (1, 2, 8):
            blocks = [shard.sample_shard(r, w, n) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def test_scale_inflate_cutoff_host_functions():
    from wisecondor_b200 import wisetools
    rng = np.random.default_rng(0)
    sample = {str(c): rng.integers(0, 50, size=int(rng.integers(1, 40))).astype(np.int32) for c in range(1, 23)}
    assert wisetools.scaleSample(sample, 5, 5) is sample and wisetools.scaleSample(sample, 5, None) is sample
    got, want = wisetools.scaleSample(sample, 2, 10), wc_oracle.scale_sample(sample, 2, 10)
    for c in sample:
        assert np.array_equal(got[c], want[c]) and got[c].dtype == np.int32
    with pytest.raises(SystemExit):
        wisetools.scaleSample(sample, 3, 10)
    mask = rng.random(50) > 0.3
    keep = rng.random(int(mask.sum())) > 0.2
    vals = rng.normal(size=int(keep.sum()))
    assert np.array_equal(wisetools.inflateArrayMulti(vals, [mask, keep]), wc_oracle.inflate_multi(vals, [mask, keep]))
    d = rng.gamma(2.0, 1.0, size=(40, 9))
    assert wisetools.getOptimalCutoff(d, 3)[0] == wc_oracle.get_optimal_cutoff(d, 3)
    counts = wisetools._refFormatCounts(sample, [len(sample[str(c)]) + (c % 3) - 1 for c in range(1, 23)])
    assert counts.dtype == np.int32 and counts.shape[0] == sum(len(sample[str(c)]) + (c % 3) - 1 for c in range(1, 23))


@pytest.mark.parametrize("variant,extra", [("default", []), ("mineffect0", ["-mineffect", "0"]), ("mineffect5", ["-mineffect", "5"])])
def test_report_prints_what_the_reference_prints(tmp_path, monkeypatch, capsys, variant, extra):
    """`report` (reference wisecondor.py:304-342) against the reference's own output on the same two npz files
    (tests/golden/report.json, written by make_golden.golden_report)."""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    import wisecondor
    make_golden.report_files(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    capsys.readouterr()
    wisecondor.main(["report", "s.npz", "o.npz"] + extra)
    got = capsys.readouterr().out
    with open(os.path.join(os.path.dirname(__file__), "golden", "report.json")) as f:
        want = json.load(f)[variant]
    assert got == want


def test_cli_surface_matches_reference():
    import wisecondor
    p = wisecondor.buildParser()
    a = p.parse_args(["newref", "a.npz", "b.npz", "out.npz"])
    assert (a.infiles, a.outfile, a.refsize, a.binsize, a.cpus, a.parts) == (["a.npz", "b.npz"], "out.npz", 100, None, 1, 1)
    assert a.func is wisecondor.toolNewref
    a = p.parse_args(["newrefpart", "p.npz", "part", "2", "5", "-refsize", "50"])
    assert a.part == [2, 5] and a.refsize == 50 and a.func is wisecondor.toolNewrefPart
    a = p.parse_args(["newrefpost", "p.npz", "part", "5", "out.npz"])
    assert a.parts == 5
    a = p.parse_args(["test", "s.npz", "o.npz", "r.npz"])
    assert (a.minzscore, a.mineffectsize, a.multitest, a.minrefbins, a.repeats) == (None, 0, 1000, 25, 5)
    assert list(a.chromosomes) == list(range(1, 23))
    a = p.parse_args(["test", "s.npz", "o.npz", "r.npz", "-chromosomes", "1,13,21", "-minzscore", "4.5"])
    assert a.chromosomes == [1, 13, 21] and a.minzscore == 4.5
    a = p.parse_args(["convert", "x.bam", "x.npz"])
    assert a.binsize == 1e6 and a.retdist == 4 and a.retthres == 4
    # the tool functions keep the reference's names: npz `arguments` pickles args.func by name
    for name in ("toolConvert", "toolNewref", "toolNewrefPrep", "toolNewrefPart", "toolNewrefPost", "toolTest",
                 "toolPlot", "toolReport"):
        assert callable(getattr(wisecondor, name))


def _gather_worker(rank, world, port, n, k, q):
    import torch
    import torch.distributed as dist
    from wisecondor_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        a, b = shard.row_shard(rank, world, n)
        rows = torch.arange(a, b, dtype=torch.int32)[:, None] * 1000 + torch.arange(k, dtype=torch.int32)[None, :]
        idx, dst = shard.allgather_rows(rows, rows.to(torch.float64) * 0.5, n)
        want = torch.arange(n, dtype=torch.int32)[:, None] * 1000 + torch.arange(k, dtype=torch.int32)[None, :]
        q.put((rank, bool(torch.equal(idx, want) and torch.equal(dst, want.to(torch.float64) * 0.5))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 101), (3, 100)])
def test_row_sharded_gather_over_gloo(world, n):
    """The N>1 newref path on CPU: every rank's getPart rows all-gathered into the whole table on every rank."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


class _CpuStandInEngine(object):
    """Test stand-in for the three device steps of a sharded symmetric search (shard._DeviceEngine): same contract on CPU
    tensors with brute-force numpy, so that the collectives and buffer layouts of SymmetricShardedSearch run under gloo.
    begin: the owned bins' thresholds; sweep: for EVERY bin its in_cap nearest candidates among this rank's bins;
    finish: merge the received candidates per owned bin and rank them (distance, index)."""
    BLOCK = 128

    def __init__(self, in_cap):
        self.in_cap = in_cap

    def dims(self, n, refsize, world, rank, device):
        nb = (n + self.BLOCK - 1) // self.BLOCK
        bp = (nb + world - 1) // world
        rows_per = bp * self.BLOCK
        return {"rows_per": rows_per, "in_cap": self.in_cap, "thr_len": world * rows_per + self.BLOCK,
                "row0": min(n, rank * rows_per), "row1": min(n, (rank + 1) * rows_per), "blocks": nb}

    def begin(self, x, chrom_bins, refsize, rank, world, thr):
        self.x, self.bins, self.k, self.rank, self.world = x.numpy(), list(chrom_bins), refsize, rank, world
        self.n = self.x.shape[0]
        self.d = self.dims(self.n, refsize, world, rank, None)
        thr.fill_(-1)
        thr[self.d["row0"]:self.d["row1"]] = -(rank + 2)            # "tighter" (smaller) than everybody else's -1
        thr[self.n:] = 0

    def sweep(self, thr, in_key, in_j, in_cnt):
        t = thr.numpy()
        for r in range(self.world):                                  # the all-reduce(MIN) delivered every owner's value
            dr = self.dims(self.n, self.k, self.world, r, None)
            assert (t[dr["row0"]:dr["row1"]] == -(r + 2)).all()
        assert (t[self.n:] == 0).all()
        chrom = np.repeat(np.arange(len(self.bins)), self.bins)
        own = np.arange(self.d["row0"], self.d["row1"])
        in_cnt.zero_()
        for j in range(self.n):
            cand = own[chrom[own] != chrom[j]]
            dist2 = ((self.x[cand] - self.x[j]) ** 2).sum(axis=1)
            order = np.lexsort((cand, dist2))[:self.in_cap]
            in_cnt[j] = len(order)
            in_key[j, :len(order)] = __import__("torch").from_numpy(dist2[order].view(np.int64).copy())
            in_j[j, :len(order)] = __import__("torch").from_numpy(cand[order].astype(np.int32))

    def finish(self, recv_key, recv_j, recv_cnt, idx, dist_out):
        chrom = np.repeat(np.arange(len(self.bins)), self.bins)
        starts = np.concatenate(([0], np.cumsum(self.bins)))
        rk, rj, rc = recv_key.numpy(), recv_j.numpy(), recv_cnt.numpy()
        rows_per = self.d["rows_per"]
        for r in range(self.d["row1"] - self.d["row0"]):
            j = self.d["row0"] + r
            ds, js = [], []
            for s in range(self.world):
                c = rc[s * rows_per + r]
                ds.append(rk[s * rows_per + r, :c].view(np.float64))
                js.append(rj[s * rows_per + r, :c])
            ds, js = np.concatenate(ds), np.concatenate(js)
            order = np.lexsort((js, ds))[:self.k]
            out_i = np.full(self.k, -1, dtype=np.int32)
            out_d = np.full(self.k, 1e10)
            sel = js[order]
            # other-chromosome coordinates (wisetools.py:386-393): bins after j's chromosome shift down by its length
            out_i[:len(order)] = np.where(sel >= starts[chrom[j] + 1], sel - self.bins[chrom[j]], sel)
            out_d[:len(order)] = ds[order]
            idx[r] = __import__("torch").from_numpy(out_i)
            dist_out[r] = __import__("torch").from_numpy(out_d)


def _sym_shard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from wisecondor_b200 import shard, synth
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        bins = [150, 90, 201, 60]                                    # 501 bins: 4 blocks of 128, the last one partial
        X = synth.corrected_like(bins, 12, seed=7)
        k = 20
        search = shard.SymmetricShardedSearch(X.shape[0], k, rank, world, torch.device("cpu"), engine=_CpuStandInEngine(32))
        search.run(torch.from_numpy(X), bins)
        idx, dst = search.gather()
        oidx, odst = wc_oracle.get_reference(X, bins, list(np.cumsum(bins)), k, 1, 1)
        ok = np.array_equal(idx.numpy(), oidx) and np.allclose(dst.numpy(), odst, rtol=1e-12, atol=0)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_symmetric_sharded_search_collectives_over_gloo(world):
    """The N>1 symmetric search on CPU: threshold all-reduce(MIN), candidate all-to-all in equal row splits, ordered
    gather - with a numpy stand-in for the three device steps (the kernels are covered by the GPU tests)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sym_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


def _sym_plan(bins, frac, world, rank, grid=148, group=2, rounds_on=1):
    """wc_debug_sym_plan (host only): tile lists and CTA schedule of the symmetric search for one rank."""
    from wisecondor_b200 import _cabi
    L = _cabi.lib()
    cb = np.ascontiguousarray(bins, dtype=np.int32)
    n = int(cb.sum())
    used = ctypes.c_longlong(0)
    buf = np.zeros(1 << 16, dtype=np.int32)
    rc = L.wc_debug_sym_plan(n, cb.ctypes.data_as(ctypes.c_void_p), len(cb), frac, world, rank, grid, group, rounds_on,
                             buf.ctypes.data_as(ctypes.c_void_p), len(buf), ctypes.byref(used))
    if rc != 0:
        buf = np.zeros(used.value, dtype=np.int32)
        _cabi.check(L.wc_debug_sym_plan(n, cb.ctypes.data_as(ctypes.c_void_p), len(cb), frac, world, rank, grid, group, rounds_on,
                                        buf.ctypes.data_as(ctypes.c_void_p), len(buf), ctypes.byref(used)))
    nb, b0, b1, na, nbb, npa, npb = (int(v) for v in buf[:7])
    nrb = b1 - b0
    pos = 8
    off_a = buf[pos:pos + nrb + 1]; pos += nrb + 1
    off_b = buf[pos:pos + nrb + 1]; pos += nrb + 1
    list_a = buf[pos:pos + na]; pos += na
    list_b = buf[pos:pos + nbb]; pos += nbb
    pieces_a = buf[pos:pos + 5 * npa].reshape(-1, 5); pos += 5 * npa
    pieces_b = buf[pos:pos + 5 * npb].reshape(-1, 5); pos += 5 * npb
    return dict(nb=nb, b0=b0, b1=b1, off_a=off_a, off_b=off_b, list_a=list_a, list_b=list_b, pieces_a=pieces_a,
                pieces_b=pieces_b)


@pytest.mark.parametrize("bins", [
    [4985, 4864, 3961, 3824, 3619, 3423, 3183, 2928, 2825, 2711, 2701, 2678, 2304, 2147, 2051, 1808, 1624, 1562, 1183, 1261, 963, 1027],
    [997, 973, 793, 765, 724, 685, 637, 586, 565, 543, 541, 536, 461, 430, 411, 362, 325, 313, 237, 253, 193, 206],
    [128 * 7, 128 * 3, 128 * 6],                 # chromosome ends on block boundaries
    [3100, 30, 20],                              # one chromosome holds nearly everything
    [0, 9, 0, 0, 14, 3, 0],                      # fewer bins than one block
    [300, 1, 299, 1, 640, 127, 129],
])
@pytest.mark.parametrize("world,frac", [(1, 8), (2, 8), (3, 4), (8, 16), (5, 2)])
def test_symmetric_plan_reaches_every_block_pair_exactly_once(bins, world, frac):
    """Host logic of the (sharded) symmetric search: over all ranks, every pair of 128-bin blocks that holds at least one
    bin pair of different chromosomes is seen by BOTH of its blocks exactly once (pass A: from each side; pass B: one
    tile serves both) - never twice, which would duplicate candidates - and every tile belongs to exactly one CTA piece."""
    n = int(sum(bins))
    nb = (n + 127) // 128
    chrom = np.repeat(np.arange(len(bins)), bins)
    sets = [frozenset(chrom[i * 128:(i + 1) * 128].tolist()) for i in range(nb)]
    seen = np.zeros((nb, nb), dtype=np.int32)          # seen[I][J]: how often block I learns about the bins of block J
    owned = np.zeros(nb, dtype=np.int32)
    for rank in range(world):
        p = _sym_plan(bins, frac, world, rank, grid=37, group=2)
        assert p["nb"] == nb
        for local in range(p["b1"] - p["b0"]):
            I = p["b0"] + local
            owned[I] += 1
            for t in p["list_a"][p["off_a"][local]:p["off_a"][local + 1]]:
                seen[I, t] += 1
            for t in p["list_b"][p["off_b"][local]:p["off_b"][local + 1]]:
                assert t != I
                seen[I, t] += 1
                seen[t, I] += 1
        for off, pieces in ((p["off_a"], p["pieces_a"]), (p["off_b"], p["pieces_b"])):
            count = [np.zeros(off[i + 1] - off[i], dtype=np.int32) for i in range(p["b1"] - p["b0"])]
            for cta, rb, q0, q1, step in pieces:
                assert 0 <= cta < 37 and q0 < q1 and step >= 1
                count[rb][q0:q1:step] += 1
            assert all((c == 1).all() for c in count), "a tile is missing from or repeated in the CTA schedule"
    assert (owned == 1).all()
    needed = np.array([[not (len(sets[i]) == 1 and sets[i] == sets[j]) for j in range(nb)] for i in range(nb)])
    assert (seen <= 1).all(), "a block pair is computed twice"
    assert (seen[needed] == 1).all(), "a block pair with candidates is never computed"


@pytest.mark.parametrize("n,world", [(501, 3), (57633, 8), (11537, 2), (128, 4), (1, 1)])
def test_shard_dims_partition_the_bins(n, world):
    """wc_newref_shard_dims (host arithmetic; NULL context): the ranks' bin ranges are consecutive, block-aligned, cover
    [0, N) exactly, and the exchange buffers have equal splits."""
    from wisecondor_b200 import _cabi
    L = _cabi.lib()
    prev_end, rows_per = 0, None
    for rank in range(world):
        out = (ctypes.c_longlong * 6)()
        _cabi.check(L.wc_newref_shard_dims(None, n, 100, world, rank, out))
        rp, in_cap, thr_len, row0, row1, nb = (int(v) for v in out)
        rows_per = rows_per or rp
        assert rp == rows_per and rp % 128 == 0 and nb == (n + 127) // 128
        assert row0 == min(n, rank * rp) == prev_end or (row0 == n and prev_end == n)
        assert row0 <= row1 <= n and (row1 == n or row1 % 128 == 0)
        assert thr_len >= world * rp + 128 and thr_len >= nb * 128
        assert in_cap >= 256 and in_cap & (in_cap - 1) == 0
        prev_end = row1
    assert prev_end == n
