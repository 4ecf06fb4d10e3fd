"""pb_copy_h2d / pb_copy_d2h throughput from pageable numpy memory vs the number of copy threads."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one():
    import torch

    from picasso_b200 import _lib
    l = _lib.load()
    vp, sz = C.c_void_p, C.c_size_t
    l.pb_copy_h2d.argtypes = [vp, vp, sz, vp]
    l.pb_copy_d2h.argtypes = [vp, vp, sz, vp]
    n = 1 << 30
    host = np.random.default_rng(0).integers(0, 255, n, dtype=np.uint8)
    back = np.empty_like(host)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for name, fn in (("h2d", lambda: l.pb_copy_h2d(dev.data_ptr(), host.ctypes.data, n, st)),
                     ("d2h", lambda: l.pb_copy_d2h(back.ctypes.data, dev.data_ptr(), n, st))):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            assert fn() == 0
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        res[name + "_GBs"] = n / best / 1e9
    assert np.array_equal(host, back)
    # destination freshly allocated (untouched pages: the copy pays the page faults)
    ts = []
    for _ in range(3):
        fresh = np.empty(n, np.uint8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        l.pb_copy_d2h(fresh.ctypes.data, dev.data_ptr(), n, st)
        ts.append(time.perf_counter() - t0)
        del fresh
    res["d2h_fresh_destination_GBs"] = n / min(ts) / 1e9
    t0 = time.perf_counter(); dev.copy_(torch.from_numpy(host)); torch.cuda.synchronize()
    res["torch_pageable_h2d_GBs"] = n / (time.perf_counter() - t0) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run_one()
    else:
        print(open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip(), flush=True)
        print("cores", os.cpu_count(), flush=True)
        for t in (4, 8, 12, 16):
            for stream in (0, 1):
                env = dict(os.environ, PB_COPY_THREADS=str(t), PB_COPY_STREAM=str(stream))
                out = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
                print(json.dumps({"threads": t, "streaming_stores": stream}), out.stdout.strip() or out.stderr[-300:],
                      flush=True)
