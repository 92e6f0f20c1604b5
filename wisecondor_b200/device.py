"""Device-resident entry points: torch tensors own the HBM buffers, the C ABI does the work.

PyTorch is plumbing here (allocation, streams, torch.distributed); every kernel lives in libwisecondor_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _cabi


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _cabi.WisecondorError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype or not t.is_contiguous():
        raise _cabi.WisecondorError("%s must be contiguous %s" % (name, dtype))


def newref_topk(corrected, chrom_bins, row_begin, row_end, refsize, out_idx=None, out_dist=None):
    """Reference-bin search for target rows [row_begin, row_end) (wisetools.py:364-398, 298-325).

    corrected: CUDA float64 [N][S] (bin-major).  Returns (indexes int32 [rows][refsize], distances float64
    [rows][refsize]) on the same device; indexes are positions in the other-chromosome concatenation.
    """
    _require_cuda(corrected, torch.float64, "corrected")
    n, s = corrected.shape
    dev = corrected.device
    cb = np.ascontiguousarray(chrom_bins, dtype=np.int32)
    rows = int(row_end) - int(row_begin)
    if out_idx is None:
        out_idx = torch.empty((max(rows, 0), refsize), dtype=torch.int32, device=dev)
    if out_dist is None:
        out_dist = torch.empty((max(rows, 0), refsize), dtype=torch.float64, device=dev)
    ctx = _cabi.context(dev.index if dev.index is not None else torch.cuda.current_device())
    rc = _cabi.lib().wc_newref_topk(ctx.handle, _ptr(corrected), n, s, cb.ctypes.data_as(ctypes.c_void_p), len(cb),
                                    int(row_begin), int(row_end), int(refsize), _ptr(out_idx), _ptr(out_dist),
                                    _stream_ptr(dev))
    _cabi.check(rc)
    return out_idx, out_dist


def newref_topk_host(corrected, chrom_bins, row_begin, row_end, refsize, device=0):
    """Same search with HOST numpy buffers: copies in and out inside the C call (wc_newref_topk_host)."""
    X = np.ascontiguousarray(corrected, dtype=np.float64)
    n, s = X.shape
    cb = np.ascontiguousarray(chrom_bins, dtype=np.int32)
    rows = int(row_end) - int(row_begin)
    idx = np.empty((max(rows, 0), refsize), dtype=np.int32)
    dist = np.empty((max(rows, 0), refsize), dtype=np.float64)
    ctx = _cabi.context(device)
    rc = _cabi.lib().wc_newref_topk_host(ctx.handle, X.ctypes.data_as(ctypes.c_void_p), n, s,
                                         cb.ctypes.data_as(ctypes.c_void_p), len(cb), int(row_begin), int(row_end),
                                         int(refsize), idx.ctypes.data_as(ctypes.c_void_p),
                                         dist.ctypes.data_as(ctypes.c_void_p))
    _cabi.check(rc)
    return idx, dist


def last_search_stats(device=0):
    """Device timings (ms) and counters of the most recent search on `device`."""
    ctx = _cabi.context(device)
    return {
        "center_norms_ms": ctx.phase_ms(0), "dist_topk_ms": ctx.phase_ms(1), "finalize_ms": ctx.phase_ms(2),
        "exhaustive_ms": ctx.phase_ms(3), "launches": ctx.counter(0), "exhaustive_rows": ctx.counter(1),
        "tiles": ctx.counter(3), "ctas": ctx.counter(4),
    }
