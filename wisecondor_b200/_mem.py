"""Device buffers of the host layer.

Two interchangeable owners of HBM buffers sit under wisecondor_b200.device:

  * torch tensors - the default for library use (streams, pinned memory, torch.distributed, interoperability);
  * DevArray      - a cudaMalloc block owned through the C ABI (wc_dev_alloc / wc_copy_*), no PyTorch involved.  The
                    command line selects it (WISECONDOR_BACKEND=native) because `import torch` alone costs seconds - more
                    than a whole `test` of one sample.

Every function of wisecondor_b200.device accepts either kind and returns the kind it was given; where nothing is given
(host numpy inputs) the module-level BACKEND decides.  Copies of the native backend are synchronous on the default stream.
"""
import ctypes
import os

import numpy as np

from . import _cabi

BACKEND = os.environ.get("WISECONDOR_BACKEND", "torch").lower()       # "torch" | "native"


class _Dev(object):
    def __init__(self, index):
        self.index = int(index)
        self.type = "cuda"


class DevArray(object):
    """A C-contiguous array in device memory: shape, numpy dtype, raw pointer.  Not a tensor library: no arithmetic, no
    views - only what the C ABI needs (a pointer) and the way back to numpy."""
    is_cuda = True

    def __init__(self, shape, dtype, device=0):
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        self.device = _Dev(device)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._ctx = _cabi.context(self.device.index)
        self._ptr = _cabi.lib().wc_dev_alloc(self._ctx.handle, self.nbytes)
        if not self._ptr:
            raise _cabi.WisecondorError("device allocation failed: %s" % _cabi.lib().wc_last_error().decode())

    def data_ptr(self):
        return self._ptr

    def is_contiguous(self):
        return True

    def numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        _cabi.check(_cabi.lib().wc_copy_d2h(self._ctx.handle, out.ctypes.data_as(ctypes.c_void_p), self._ptr, self.nbytes))
        return out

    @classmethod
    def from_numpy(cls, array, device=0):
        a = np.ascontiguousarray(array)
        out = cls(a.shape, a.dtype, device)
        _cabi.check(_cabi.lib().wc_copy_h2d(out._ctx.handle, out._ptr, a.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def __del__(self):
        try:
            if getattr(self, "_ptr", None) and self._ctx.handle:
                _cabi.lib().wc_dev_free(self._ctx.handle, self._ptr)
                self._ptr = None
        except Exception:
            pass


_TORCH = None


def torch():
    """PyTorch, imported on first use only."""
    global _TORCH
    if _TORCH is None:
        import torch as t
        _TORCH = t
    return _TORCH


def is_native(x):
    return isinstance(x, DevArray)


def is_device(x):
    return isinstance(x, DevArray) or (hasattr(x, "is_cuda") and not isinstance(x, np.ndarray) and bool(x.is_cuda))


def _torch_dtype(dtype):
    t = torch()
    return {np.dtype(np.float64): t.float64, np.dtype(np.int32): t.int32, np.dtype(np.uint8): t.uint8,
            np.dtype(np.int64): t.int64}[np.dtype(dtype)]


def np_dtype(x):
    if isinstance(x, DevArray):
        return x.dtype
    t = torch()
    return {t.float64: np.dtype(np.float64), t.int32: np.dtype(np.int32), t.uint8: np.dtype(np.uint8),
            t.int64: np.dtype(np.int64)}.get(x.dtype)


def device_index(x):
    if isinstance(x, DevArray):
        return x.device.index
    return x.device.index if x.device.index is not None else torch().cuda.current_device()


def empty(shape, dtype, like=None, device=None):
    """Uninitialised device array of the same kind as `like` (or of the BACKEND's kind) on `like`'s device."""
    native = is_native(like) if like is not None else BACKEND == "native"
    dev = device_index(like) if like is not None else int(device or 0)
    if native:
        return DevArray(shape, dtype, dev)
    t = torch()
    return t.empty(tuple(shape), dtype=_torch_dtype(dtype), device=t.device("cuda", dev))


def to_device(array, device=0, like=None):
    """Host numpy array -> device array (kind as for `empty`)."""
    native = is_native(like) if like is not None else BACKEND == "native"
    dev = device_index(like) if like is not None else int(device or 0)
    if native:
        return DevArray.from_numpy(array, dev)
    t = torch()
    return t.as_tensor(np.ascontiguousarray(array), device=t.device("cuda", dev))


def to_host(x):
    """Device array -> numpy (synchronises)."""
    if isinstance(x, DevArray):
        return x.numpy()
    return x.cpu().numpy()


def synchronize(x):
    """Wait for the work enqueued on x's stream."""
    if isinstance(x, DevArray):
        _cabi.check(_cabi.lib().wc_dev_sync(x._ctx.handle))
    else:
        torch().cuda.current_stream(x.device).synchronize()


def ptr(x):
    return ctypes.c_void_p(x.data_ptr()) if x is not None else None


def stream_ptr(x):
    """The stream work on `x` is enqueued on: torch's current stream of x's device, or the default stream."""
    if isinstance(x, DevArray):
        return None
    return ctypes.c_void_p(torch().cuda.current_stream(x.device).cuda_stream)


def require(x, dtype, name):
    if not is_device(x):
        raise _cabi.WisecondorError("%s must live in device memory (there is no CPU path)" % name)
    if np_dtype(x) != np.dtype(dtype) or not x.is_contiguous():
        raise _cabi.WisecondorError("%s must be contiguous %s" % (name, np.dtype(dtype)))
