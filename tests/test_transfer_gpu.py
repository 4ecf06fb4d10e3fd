"""pb_copy_h2d / pb_copy_d2h (csrc/transfer.cu): pageable numpy memory through the threaded pinned
staging, pinned memory directly -- byte-exact round trips for sizes below, at and above the
staging chunk, with odd lengths."""
import ctypes as C

import numpy as np
import pytest

from picasso_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbytes", [1, 4097, (1 << 20) - 3, (1 << 20) + 5, (32 << 20) + 12345, (70 << 20) + 1])
def test_pageable_round_trip(nbytes):
    import torch

    l = _lib.load()
    vp, sz = C.c_void_p, C.c_size_t
    l.pb_copy_h2d.argtypes = [vp, vp, sz, vp]
    l.pb_copy_d2h.argtypes = [vp, vp, sz, vp]
    rng = np.random.default_rng(nbytes)
    src = rng.integers(0, 256, nbytes, dtype=np.uint8)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(l.pb_copy_h2d(dev.data_ptr(), src.ctypes.data, nbytes, st))
    keep = src.copy()
    src[:] = 0                                   # a pageable source may be reused right after the call
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), keep)
    back = np.empty(nbytes, np.uint8)
    _lib.check(l.pb_copy_d2h(back.ctypes.data, dev.data_ptr(), nbytes, st))
    assert np.array_equal(back, keep)             # blocking: complete on return


def test_pinned_round_trip_and_errors():
    import torch

    l = _lib.load()
    vp, sz = C.c_void_p, C.c_size_t
    l.pb_copy_h2d.argtypes = [vp, vp, sz, vp]
    l.pb_copy_d2h.argtypes = [vp, vp, sz, vp]
    n = (8 << 20) + 7
    pin = _lib.PinnedArray((n,), np.uint8)
    pin.array[:] = np.random.default_rng(1).integers(0, 256, n, dtype=np.uint8)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(l.pb_copy_h2d(dev.data_ptr(), pin.array.ctypes.data, n, st))
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), pin.array)
    out = _lib.PinnedArray((n,), np.uint8)
    _lib.check(l.pb_copy_d2h(out.array.ctypes.data, dev.data_ptr(), n, st))
    assert np.array_equal(out.array, pin.array)
    assert l.pb_copy_h2d(None, pin.array.ctypes.data, n, st) == 1          # PB_ERR_INVALID
    assert l.pb_copy_d2h(out.array.ctypes.data, dev.data_ptr(), 0, st) == 0
    pin.free(); out.free()
