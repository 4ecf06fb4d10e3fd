"""Where the time of the Python-facing calls goes (run on the GPU box):
  * pb_copy_h2d of 1.96 GB from pageable memory alone
  * pb_mle_fit (C ABI) with pageable input + page-locked outputs, call by call
  * picasso_b200.gaussmle.gaussmle(pageable ndarray), call by call, results released between calls
  * zfit.zfit on 10 M localizations (fused device table path), lib.ensure_sanity, lib.locs_to_records
One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def alloc_probe():
    """pb_host_alloc of 560 MB: time and DMA rate from the block (run once per PB_HOST_HUGEPAGES mode)."""
    import ctypes as C

    import torch

    from picasso_b200 import _lib

    l = _lib.load()
    nbytes = 560_000_000
    torch.zeros(1, device="cuda")
    res = {"hugepages": os.environ.get("PB_HOST_HUGEPAGES", "default(1)")}
    ts = []
    ptrs = []
    for _ in range(3):
        p = C.c_void_p()
        t0 = time.perf_counter()
        _lib.check(l.pb_host_alloc(C.byref(p), nbytes))
        ts.append(time.perf_counter() - t0)
        ptrs.append(p)
    res["alloc_ms"] = [round(1e3 * t, 1) for t in ts]
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for name, fn in (("h2d", lambda: l.pb_copy_h2d(dev.data_ptr(), ptrs[0], nbytes, st)),
                     ("d2h", lambda: l.pb_copy_d2h(ptrs[0], dev.data_ptr(), nbytes, st))):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        res[name + "_GBs"] = round(nbytes / best / 1e9, 1)
    t0 = time.perf_counter()
    for p in ptrs:
        l.pb_host_free(p)
    res["free_ms_each"] = round(1e3 * (time.perf_counter() - t0) / 3, 1)
    print(json.dumps(res), flush=True)


def main():
    import torch

    import bench
    from picasso_b200 import _lib, gaussmle, lib as pblib, testing, zfit

    l = _lib.load()
    dev = torch.device("cuda", 0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    out = {"n": n, "copy_threads": os.environ.get("PB_COPY_THREADS"), "stream": os.environ.get("PB_COPY_STREAM")}
    spots = bench.gen_spots_device(torch, n, 7, 1000, dev)
    hp = spots.cpu().numpy()
    dbuf = torch.empty_like(spots)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(l.pb_copy_h2d(dbuf.data_ptr(), hp.ctypes.data, hp.nbytes, st)); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    out["h2d_pageable_ms"] = [round(1e3 * t, 2) for t in ts]
    del dbuf
    th = _lib.pinned_empty((n, 6), np.float32); cr = _lib.pinned_empty((n, 6), np.float32)
    ll = _lib.pinned_empty((n,), np.float32); it = _lib.pinned_empty((n,), np.int32)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        _lib.check(l.pb_mle_fit(n, 7, _lib.ptr(hp), 0.001, 100, 1, _lib.ptr(th), _lib.ptr(cr), _lib.ptr(ll),
                                _lib.ptr(it), None, None))
        ts.append(time.perf_counter() - t0)
    out["pb_mle_fit_pageable_in_pinned_out_ms"] = [round(1e3 * t, 2) for t in ts]
    del th, cr, ll, it
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        r = gaussmle.gaussmle(hp, 0.001, 100, "sigmaxy")
        ts.append(time.perf_counter() - t0)
        del r
    out["gaussmle_python_ms"] = [round(1e3 * t, 2) for t in ts]
    out["gaussmle_python_Mfits_per_s_best"] = n / min(ts) / 1e6
    ts = []
    held = None
    for _ in range(5):                       # the caller keeps the previous result alive (bench.py's loop)
        t0 = time.perf_counter()
        held = gaussmle.gaussmle(hp, 0.001, 100, "sigmaxy")
        ts.append(time.perf_counter() - t0)
    out["gaussmle_python_results_held_ms"] = [round(1e3 * t, 2) for t in ts]
    del held, hp, spots
    torch.cuda.empty_cache()
    # ---- z fit / sanity / records on 10 M localizations ----
    locs, info, calib = testing.synthetic_zfit_locs(n, 5)
    for flt in (2, 0):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            res, _ = zfit.zfit(locs, list(info), calibration=dict(calib), fitting_method="gaussmle", filter=flt)
            ts.append(time.perf_counter() - t0)
        out[f"zfit_api_filter{flt}_s"] = [round(t, 4) for t in ts]
        out[f"zfit_rows_kept_filter{flt}"] = len(res)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); s = pblib.ensure_sanity(res, info); ts.append(time.perf_counter() - t0)
    out["ensure_sanity_s"] = [round(t, 4) for t in ts]
    t0 = time.perf_counter(); h = pblib._ensure_sanity_host(res, info); out["ensure_sanity_host_numpy_s"] = round(time.perf_counter() - t0, 4)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); rec = pblib.locs_to_records(res, info); ts.append(time.perf_counter() - t0)
    out["locs_to_records_s"] = [round(t, 4) for t in ts]
    t0 = time.perf_counter(); ref = h.to_records(index=False); out["to_records_pandas_s"] = round(time.perf_counter() - t0, 4)
    out["records_equal"] = bool(rec.tobytes() == ref.tobytes())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "alloc":
        alloc_probe()
    else:
        main()
