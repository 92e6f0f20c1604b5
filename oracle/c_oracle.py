"""TEST INFRASTRUCTURE - ctypes loader for oracle/libwc_oracle.so (the C restatement in wc_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libwc_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.wc_oracle_get_reference.restype = ctypes.c_int
        L.wc_oracle_get_reference.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.wc_oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def get_reference_rows(corrected, chrom_bins, row_begin, row_end, k, nthreads=0):
    """Exact search for rows [row_begin, row_end): returns (indexes int32 rows x k, distances f64 rows x k)."""
    X = np.ascontiguousarray(corrected, dtype=np.float64)
    n, s = X.shape
    cb = np.ascontiguousarray(chrom_bins, dtype=np.int32)
    rows = row_end - row_begin
    idx = np.empty((rows, k), dtype=np.int32)
    dst = np.empty((rows, k), dtype=np.float64)
    rc = lib().wc_oracle_get_reference(X.ctypes.data, n, s, cb.ctypes.data, len(cb), row_begin, row_end, k,
                                       idx.ctypes.data, dst.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("wc_oracle_get_reference failed: %d" % rc)
    return idx, dst


def max_threads():
    return lib().wc_oracle_max_threads()
