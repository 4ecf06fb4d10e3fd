"""Time pb_mle_fit (host buffers, pinned) for several chunk sizes (PB_MLE_CHUNK_MB)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from picasso_b200 import _lib

lib = _lib.load()
n = 10_000_000
spots = bench.gen_spots_device(torch, n, 7, 1000, torch.device("cuda", 0))
hs = _lib.PinnedArray((n, 7, 7), np.float32); torch.from_numpy(hs.array).copy_(spots.cpu())
hth = _lib.PinnedArray((n, 6), np.float32); hcr = _lib.PinnedArray((n, 6), np.float32)
hll = _lib.PinnedArray((n,), np.float32); hit = _lib.PinnedArray((n,), np.int32)
del spots
for mb in (8, 16, 32, 64, 128, 256):
    os.environ["PB_MLE_CHUNK_MB"] = str(mb)
    def go():
        _lib.check(lib.pb_mle_fit(n, 7, _lib.ptr(hs.array), 0.001, 100, 1, _lib.ptr(hth.array),
                                  _lib.ptr(hcr.array), _lib.ptr(hll.array), _lib.ptr(hit.array), None, None))
    go(); go()
    t0 = time.perf_counter()
    for _ in range(3):
        go()
    el = (time.perf_counter() - t0) / 3
    print(json.dumps({"chunk_mb": mb, "ms": el * 1e3, "Mfits_per_s": n / el / 1e6,
                      "GBps_h2d_plus_d2h": 2.52 / el}), flush=True)
