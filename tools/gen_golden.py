"""Generate the golden vectors under tests/golden/ by running the REAL reference
(jungmannlab/picasso mounted at /root/reference, imported with its GUI/IO
dependencies mocked -- tools/ref_import.py).

Run in the build container only:

    python tools/gen_golden.py [mle] [identify] [lq] [render] [undrift] [testdata]

Each fixture stores the inputs and the reference's outputs, so the tests need
neither the reference nor this script at run time.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_mle(ref):
    from picasso_b200 import testing

    gm = ref["gaussmle"]
    # config 1: 10k Poisson spots, box 7 -- counts fit in uint16
    spots = testing.synthetic_spots(10_000, 7, seed=0)
    assert spots.max() < 65535 and (spots == np.round(spots)).all()
    out = {"spots_u16": spots.astype(np.uint16)}
    for method in ("sigmaxy", "sigma"):
        th, cr, ll, it = gm.gaussmle(spots, 0.001, 100, method)
        out[f"{method}_thetas"] = th
        out[f"{method}_crlbs"] = cr
        out[f"{method}_logliks"] = ll
        out[f"{method}_iterations"] = it
    save("mle_config1.npz", **out)

    # other box sizes, 300 spots each, both methods
    out = {}
    for box in (5, 9, 11, 13, 15):
        spots = testing.synthetic_spots(300, box, seed=100 + box)
        out[f"b{box}_spots_u16"] = spots.astype(np.uint16)
        for method in ("sigmaxy", "sigma"):
            th, cr, ll, it = gm.gaussmle(spots, 0.001, 100, method)
            out[f"b{box}_{method}_thetas"] = th
            out[f"b{box}_{method}_crlbs"] = cr
            out[f"b{box}_{method}_logliks"] = ll
            out[f"b{box}_{method}_iterations"] = it
    save("mle_boxes.npz", **out)

    # non-integer (noiseless, point-sampled) float spots + tighter eps / capped
    # iterations: the regimes the reference's own tests use
    # (tests/conftest.py:121-188, tests/test_gaussmle.py:129-140)
    rng = np.random.default_rng(42)
    n, box = 64, 7
    half = box // 2
    grid = np.arange(-half, half + 1, dtype=np.float64)
    x0 = rng.uniform(-0.5, 0.5, n); y0 = rng.uniform(-0.5, 0.5, n)
    sx = rng.uniform(0.8, 1.4, n); sy = rng.uniform(0.8, 1.4, n)
    ph = rng.uniform(2000, 8000, n); bg = rng.uniform(5, 50, n)
    gx = np.exp(-0.5 * ((grid[None] - x0[:, None]) / sx[:, None]) ** 2) / (sx[:, None] * np.sqrt(2 * np.pi))
    gy = np.exp(-0.5 * ((grid[None] - y0[:, None]) / sy[:, None]) ** 2) / (sy[:, None] * np.sqrt(2 * np.pi))
    clean = (ph[:, None, None] * gy[:, :, None] * gx[:, None, :] + bg[:, None, None]).astype(np.float32)
    out = {"spots": clean}
    for tag, eps, max_it in (("e3", 1e-3, 100), ("e6", 1e-6, 100), ("it3", 1e-3, 3), ("it0", 1e-3, 0)):
        for method in ("sigmaxy", "sigma"):
            th, cr, ll, it = gm.gaussmle(clean, eps, max_it, method)
            out[f"{tag}_{method}_thetas"] = th
            out[f"{tag}_{method}_crlbs"] = cr
            out[f"{tag}_{method}_logliks"] = ll
            out[f"{tag}_{method}_iterations"] = it
    save("mle_float_spots.npz", **out)


GENERATORS = {"mle": gen_mle}


def main():
    which = sys.argv[1:] or list(GENERATORS)
    ref = ref_import.import_reference()
    for w in which:
        GENERATORS[w](ref)


if __name__ == "__main__":
    main()
