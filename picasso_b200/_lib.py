"""ctypes binding of libpicasso_b200.so (the C ABI in include/picasso_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or no B200 is
visible, the product functions raise.  (The CPU oracle under ``oracle/`` is test
infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpicasso_b200.so")

_lib = None


class PicassoB200Error(RuntimeError):
    """Raised when a libpicasso_b200 call returns a non-zero status."""


def _declare(lib):
    vp, i32, i64, f64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_size_t
    lib.pb_last_error.restype = C.c_char_p
    lib.pb_version.restype = C.c_char_p
    lib.pb_device_count.restype = i32
    lib.pb_launch_count.restype = i64
    lib.pb_set_device.argtypes = [i32]
    lib.pb_get_device.argtypes = [C.POINTER(i32)]
    lib.pb_get_device.restype = i32
    lib.pb_host_alloc.argtypes = [C.POINTER(vp), sz]
    lib.pb_host_free.argtypes = [vp]
    lib.pb_mle_fit.argtypes = [sz, i32, vp, f64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pb_mle_fit_dev.argtypes = [sz, i32, vp, f64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pb_mle_set_impl.argtypes = [i32]
    lib.pb_mle_profile.argtypes = [i32]
    lib.pb_mle_profile_read.argtypes = [vp]
    for name in ("pb_set_device", "pb_synchronize", "pb_host_alloc", "pb_host_free",
                 "pb_mle_fit", "pb_mle_fit_dev", "pb_mle_set_impl", "pb_mle_get_impl",
                 "pb_mle_profile", "pb_mle_profile_read"):
        getattr(lib, name).restype = i32


def load():
    """Load the shared library (once). Raises ImportError if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m picasso_b200.build` "
                "(nvcc, sm_100a). picasso_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = load().pb_last_error().decode("utf-8", "replace")
        if status == 1:
            raise ValueError(msg)
        raise PicassoB200Error(f"status = {status}, message = {msg}")


def device_count() -> int:
    return int(load().pb_device_count())


def require_gpu() -> None:
    if device_count() < 1:
        raise PicassoB200Error(
            "no sm_100 (B200) device visible: picasso_b200 runs its hot path on the GPU only"
        )


def current_device() -> int:
    """CUDA device of the calling thread."""
    d = C.c_int(0)
    check(load().pb_get_device(C.byref(d)))
    return int(d.value)


def on_callers_device(fn):
    """Wrap ``fn`` for execution on another host thread: the CUDA current device is per thread,
    so a worker spawned by a rank that selected GPU r (pb_set_device / torch.cuda.set_device) would
    otherwise run on device 0.  The caller's device is captured now and selected in the worker."""
    dev = current_device() if device_count() > 0 else None

    def run(*a, **k):
        if dev is not None:
            check(load().pb_set_device(dev))
        return fn(*a, **k)

    return run


def ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def launch_count() -> int:
    return int(load().pb_launch_count())


class PinnedArray:
    """A numpy array backed by pinned (page-locked) host memory from pb_host_alloc."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        check(load().pb_host_alloc(C.byref(p), n))
        self._p = p
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p is not None:
            self.array = None
            load().pb_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
