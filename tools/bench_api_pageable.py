"""User-facing Python API timings with ordinary (pageable) numpy arrays -- what a picasso user
passes -- next to pinned buffers.  Run on the GPU box: python tools/bench_api_pageable.py"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from picasso_b200 import _lib, gausslq, gaussmle, render, testing  # noqa: E402


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    base = testing.synthetic_spots(200_000, 7, seed=1)
    spots = np.tile(base, (n // len(base), 1, 1))
    gaussmle.gaussmle(spots[:10000], 1e-3, 100, "sigmaxy")
    out = {"n_spots": len(spots), "spots_bytes": int(spots.nbytes)}
    t = best(lambda: gaussmle.gaussmle(spots, 1e-3, 100, "sigmaxy"))
    out["gaussmle_pageable"] = {"seconds": t, "fits_per_s": len(spots) / t, "GBs_in": spots.nbytes / t / 1e9}
    pin = _lib.PinnedArray(spots.shape, spots.dtype)
    pin.array[...] = spots
    t = best(lambda: gaussmle.gaussmle(pin.array, 1e-3, 100, "sigmaxy"))
    out["gaussmle_pinned_input"] = {"seconds": t, "fits_per_s": len(spots) / t}
    m = min(n, 2_000_000)
    gausslq.fit_spots(spots[:10000])
    t = best(lambda: gausslq.fit_spots(spots[:m]))
    out["gausslq_pageable"] = {"n": m, "seconds": t, "fits_per_s": m / t}
    pin.free()
    import pandas as pd
    rng = np.random.default_rng(2)
    k = 10_000_000
    locs = pd.DataFrame({"x": rng.uniform(0, 512, k).astype(np.float32), "y": rng.uniform(0, 512, k).astype(np.float32),
                         "lpx": rng.uniform(0.02, 0.08, k).astype(np.float32),
                         "lpy": rng.uniform(0.02, 0.08, k).astype(np.float32)})
    info = [{"Height": 512, "Width": 512, "Frames": 1, "Pixelsize": 130}]
    render.render(locs.iloc[:1000], info, oversampling=20, blur_method="gaussian")
    t = best(lambda: render.render(locs, info, oversampling=20, blur_method="gaussian"))
    out["render_gaussian_10M_os20"] = {"seconds": t, "locs_per_s": k / t}
    t = best(lambda: render.render(locs, info, oversampling=20, blur_method=None))
    out["render_hist_10M_os20"] = {"seconds": t, "locs_per_s": k / t}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
