"""LQ parity statistics, CUDA vs the CPU oracle (bit-identical to scipy.optimize.leastsq):
same-nfev fraction, bit-identical rows and the all-spot RMS per parameter, for both kernel
variants (pb_lq_set_impl 0 = register-resident factorisation, 1 = MINPACK-order QR), with the
kernel time of each.  Run on the GPU box; one JSON line per (box, variant), kept under profiles/.

    python tools/parity_lq.py [n_spots]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle  # noqa: E402
from picasso_b200 import _lib, gausslq, testing  # noqa: E402
from sim_lq import stats  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    lib = _lib.load()
    lib.pb_lq_set_impl.argtypes = [C.c_int]
    for box in (7, 5, 9, 13):
        m = n if box == 7 else n // 10
        spots = testing.synthetic_spots(m, box, seed=77)
        oth, oinfo, onfev = oracle.fit_spots_lq(spots, nthreads=os.cpu_count(), return_info=True)
        for impl in (0, 1):
            _lib.check(lib.pb_lq_set_impl(impl))
            gausslq._fit(spots[:1000])
            t0 = time.perf_counter()
            th, info, nfev = gausslq._fit(spots, want_info=True)
            dt = time.perf_counter() - t0
            out = {"n": m, "box": box, "impl": impl, "api_seconds": dt}
            out.update(stats(th, nfev, oth, onfev))
            out["info_equal"] = float((info == oinfo).mean())
            print(json.dumps(out), flush=True)
        _lib.check(lib.pb_lq_set_impl(0))


if __name__ == "__main__":
    main()
