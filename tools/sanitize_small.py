"""Small invocations of the kernels changed in round 2 for compute-sanitizer runs on the GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py

MLE fit (both methods, boxes 5 / 7 / 13, degenerate ROIs included) and the binned render with both
accumulation passes.  Prints one line per case; any sanitizer finding fails the run."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picasso_b200 import _lib, gaussmle, render as pbrender, testing  # noqa: E402


def main():
    lib = _lib.load()
    _lib.require_gpu()
    for box in (5, 7, 13):
        spots = testing.synthetic_spots(1500, box, seed=box)
        spots[:8] = 0.0
        spots[8] = -3.0
        spots[9, box // 2, box // 2] = 1e6
        spots[10] = np.nan
        for method in ("sigmaxy", "sigma"):
            th, cr, ll, it = gaussmle.gaussmle(spots, 0.001, 100, method)
            print("mle", box, method, int(it.sum()), flush=True)
    rng = np.random.default_rng(1)
    n = 80_000
    lp = rng.uniform(0.02, 0.08, (2, n)).astype(np.float32)
    lp[:, :500] = rng.uniform(0.15, 0.45, (2, 500)).astype(np.float32)
    locs = pd.DataFrame({"x": rng.uniform(0, 32, n).astype(np.float32), "y": rng.uniform(0, 32, n).astype(np.float32),
                         "lpx": lp[0], "lpy": lp[1]})
    info = [{"Height": 32, "Width": 32, "Frames": 1, "Pixelsize": 130}]
    imgs = []
    for impl in (0, 1):
        _lib.check(lib.pb_render_set_impl(impl))
        k, img = pbrender.render(locs, info, oversampling=20, blur_method="gaussian")
        imgs.append(img)
        print("render impl", impl, k, float(img.sum()), flush=True)
    lib.pb_render_set_impl(0)
    print("render passes agree", bool(np.allclose(imgs[0], imgs[1], rtol=1e-4, atol=1e-6 * imgs[0].max())))


if __name__ == "__main__":
    main()
