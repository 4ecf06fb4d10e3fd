"""Print one JSON line of MLE parity statistics (CUDA vs CPU oracle) -- run on
the GPU box; summaries are kept under profiles/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from picasso_b200 import gaussmle, testing  # noqa: E402


def report(n, box, method, seed=0):
    spots = testing.synthetic_spots(n, box, seed=seed)
    th, cr, ll, it = gaussmle.gaussmle(spots, 0.001, 100, method)
    oth, ocr, oll, oit = oracle.gaussmle(spots, 0.001, 100, method, nthreads=os.cpu_count())
    d = th.astype(np.float64) - oth
    same = it == oit
    out = {
        "n": n, "box": box, "method": method,
        "iteration_match": float(same.mean()),
        "mean_iterations": float(oit.mean()),
        "rms_abs": dict(zip("x y photons bg sx sy".split(), np.sqrt((d ** 2).mean(0)).tolist())),
        "rms_rel_photons_bg": np.sqrt(((d[:, 2:4] / oth[:, 2:4]) ** 2).mean(0)).tolist(),
        "max_abs_same_iter": np.abs(d[same]).max(0).tolist(),
        "theta_bit_identical_rows": float((th.view(np.uint32) == oth.view(np.uint32)).all(1).mean()),
        "crlb_max_rel_same_iter": float(np.nanmax((np.abs(cr - ocr) / np.where(ocr != 0, np.abs(ocr), np.nan))[same])),
        "crlb_zero_pattern_equal": bool((((cr == 0) == (ocr == 0))[same]).all()),
        "n_singular_fisher": int((ocr == 0).any(1).sum()),
        "loglik_max_abs_same_iter": float(np.abs(ll[same] - oll[same]).max()),
    }
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    for box, method in ((7, "sigmaxy"), (7, "sigma"), (9, "sigmaxy"), (13, "sigmaxy"), (13, "sigma")):
        print(json.dumps(report(n if box == 7 else n // 10, box, method)), flush=True)
