"""One process driving two GPUs (pb_set_device): pageable transfers, host-buffer fits and the
async entry points on device 0, then device 1, then both from two threads at once.  Needs 2
visible B200s (skips otherwise); on the single-GPU box the same code paths run on device 0."""
import threading

import numpy as np
import pytest

from picasso_b200 import _lib, gaussmle, gausslq, testing

pytestmark = pytest.mark.gpu


def _fit_on(dev, spots, out):
    lib = _lib.load()
    _lib.check(lib.pb_set_device(dev))
    out[dev] = (gaussmle.gaussmle(spots, 0.001, 100), gausslq.fit_spots(spots), _lib.current_device())


def test_two_devices_one_process():
    lib = _lib.load()
    ndev = _lib.device_count()
    devs = [0, 1] if ndev >= 2 else [0, 0]
    spots = testing.synthetic_spots(30_000, 7, seed=21)        # 5.9 MB: staged (pageable) transfers
    base = gaussmle.gaussmle(spots, 0.001, 100)
    base_lq = gausslq.fit_spots(spots)
    try:
        for d in devs:                                          # sequentially: 0, then 1
            out = {}
            _fit_on(d, spots, out)
            (th, cr, ll, it), lq, cur = out[d]
            assert cur == d
            np.testing.assert_array_equal(th, base[0]); np.testing.assert_array_equal(it, base[3])
            np.testing.assert_array_equal(lq, base_lq)
        if ndev >= 2:                                           # both at once from two threads
            out = {}
            ts = [threading.Thread(target=_fit_on, args=(d, spots, out)) for d in (0, 1)]
            [t.start() for t in ts]; [t.join() for t in ts]
            for d in (0, 1):
                np.testing.assert_array_equal(out[d][0][0], base[0])
                np.testing.assert_array_equal(out[d][1], base_lq)
    finally:
        _lib.check(lib.pb_set_device(0))


def test_async_entry_points_run_on_callers_device():
    """gaussmle_async / fit_spots_parallel(asynch=True) spawn host threads; the CUDA current device
    is per thread, so the worker must select the caller's device."""
    import time

    lib = _lib.load()
    dev = 1 if _lib.device_count() >= 2 else 0
    seen = []
    _lib.check(lib.pb_set_device(dev))
    try:
        wrapped2 = _lib.on_callers_device(lambda: seen.append(_lib.current_device()))
        t = threading.Thread(target=wrapped2); t.start(); t.join()
        assert seen == [dev]
        spots = testing.synthetic_spots(4000, 7, seed=4)
        cur, th, cr, ll, it = gaussmle.gaussmle_async(spots, 0.001, 100)
        t0 = time.time()
        while cur[0] < len(spots) and time.time() - t0 < 60:
            time.sleep(0.01)
        ref = gaussmle.gaussmle(spots, 0.001, 100)
        np.testing.assert_array_equal(th, ref[0])
        fs = gausslq.fit_spots_parallel(spots, asynch=True)
        np.testing.assert_array_equal(gausslq.fits_from_futures(fs), gausslq.fit_spots(spots))
    finally:
        _lib.check(lib.pb_set_device(0))
