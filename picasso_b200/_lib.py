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
    # plumbing used from several modules: declared here once (an undeclared ctypes function would
    # truncate 64-bit pointers passed as Python ints)
    lib.pb_copy_h2d.argtypes = [vp, vp, sz, vp]
    lib.pb_copy_d2h.argtypes = [vp, vp, sz, vp]
    lib.pb_dev_alloc.argtypes = [C.POINTER(vp), sz]
    lib.pb_dev_free.argtypes = [vp]
    lib.pb_copy_d2d_async.argtypes = [vp, vp, sz, vp]
    for name in ("pb_copy_h2d", "pb_copy_d2h", "pb_dev_alloc", "pb_dev_free", "pb_copy_d2d_async"):
        getattr(lib, name).restype = i32
    lib.pb_mle_fit.argtypes = [sz, i32, vp, f64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pb_mle_fit_dev.argtypes = [sz, i32, vp, f64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pb_mle_set_impl.argtypes = [i32]
    lib.pb_render_set_impl.argtypes = [i32]
    lib.pb_render_set_impl.restype = i32
    lib.pb_render_get_impl.restype = i32
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


# ---- pooled pinned output arrays -------------------------------------------------------------
# The arrays the Python layer RETURNS (thetas, CRLBs, images, column blocks) are allocated by this
# layer, so they can live in page-locked memory: the device -> host copy is then one direct DMA at
# PCIe speed instead of a staged copy into freshly mapped pageable pages (first-touch page faults
# cap that at ~25 GB/s).  cudaHostAlloc itself is slow (page pinning, ~0.2 s / GB), so blocks are
# recycled through a small pool: when the last numpy view of a block is garbage collected the
# block goes back to the pool (bounded by PB_PINNED_POOL_MB, default 4096) instead of being freed.
import threading as _threading

_POOL_LOCK = _threading.Lock()
_POOL = []                       # [(nbytes, ptr)] free blocks
_POOL_BYTES = 0
_POOL_LIMIT = int(os.environ.get("PB_PINNED_POOL_MB", "4096")) << 20
_PINNED_MAX = int(os.environ.get("PB_PINNED_MAX_MB", "8192")) << 20   # larger outputs stay pageable


class _PinnedBlock:
    """Owner of one page-locked block; numpy views keep it alive through ``__array_interface__``."""

    __slots__ = ("ptr", "nbytes", "__array_interface__", "__weakref__")

    def __init__(self, nbytes: int):
        global _POOL_BYTES
        self.ptr = None
        with _POOL_LOCK:
            best = None
            for k, (sz, _) in enumerate(_POOL):
                if sz >= nbytes and sz <= max(nbytes + (nbytes >> 2), nbytes + (1 << 20)):
                    if best is None or sz < _POOL[best][0]:
                        best = k
            if best is not None:
                sz, p = _POOL.pop(best)
                _POOL_BYTES -= sz
                self.ptr, self.nbytes = p, sz
        if self.ptr is None:
            p = C.c_void_p()
            check(load().pb_host_alloc(C.byref(p), max(nbytes, 1)))
            self.ptr, self.nbytes = p.value, max(nbytes, 1)
        self.__array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1",
                                    "data": (self.ptr, False), "version": 3}

    def __del__(self):
        global _POOL_BYTES
        try:
            p, self.ptr = self.ptr, None
            if p is None:
                return
            with _POOL_LOCK:
                if _POOL_BYTES + self.nbytes <= _POOL_LIMIT:
                    _POOL.append((self.nbytes, p))
                    _POOL_BYTES += self.nbytes
                    return
            load().pb_host_free(C.c_void_p(p))
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """``np.empty(shape, dtype)`` backed by pooled page-locked memory (plain ``np.empty`` for empty or
    very large arrays, or when PB_PINNED_OUTPUTS=0).  The block returns to the pool when the array
    and all its views are gone."""
    dtype = np.dtype(dtype)
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    nbytes = n * dtype.itemsize
    if nbytes == 0 or nbytes > _PINNED_MAX or os.environ.get("PB_PINNED_OUTPUTS", "1") == "0":
        return np.empty(shape, dtype)
    blk = _PinnedBlock(nbytes)
    return np.asarray(blk)[:nbytes].view(dtype).reshape(shape)


def pinned_pool_trim() -> None:
    """Free every cached block of the pinned pool."""
    global _POOL_BYTES
    with _POOL_LOCK:
        blocks, _POOL[:] = list(_POOL), []
        _POOL_BYTES = 0
    for _, p in blocks:
        load().pb_host_free(C.c_void_p(p))


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
