"""Compare the host build of the CUDA least-squares optimiser (tests/host_sim/lq_sim.cpp =
picasso_b200/csrc/lq_core.cuh compiled with g++) with the oracle (bit-identical to scipy's
leastsq): same-nfev fraction, all-spot RMS per parameter.  No GPU needed.

    python tools/sim_lq.py [n_spots] [box]
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIM_DIR = os.path.join(ROOT, "tests", "host_sim")
SIM_LIB = os.path.join(SIM_DIR, "liblq_sim.so")


def build():
    src = os.path.join(SIM_DIR, "lq_sim.cpp")
    core = os.path.join(ROOT, "picasso_b200", "csrc", "lq_core.cuh")
    if (not os.path.exists(SIM_LIB)
            or os.path.getmtime(SIM_LIB) < max(os.path.getmtime(src), os.path.getmtime(core))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC",
                               "-shared", "-x", "c++", src, "-o", SIM_LIB])
    lib = C.CDLL(SIM_LIB)
    vp = C.c_void_p
    lib.sim_lq.argtypes = [vp, C.c_longlong, C.c_int, C.c_int, vp, vp, vp]
    return lib


def sim(spots, variant=0):
    lib = build()
    sp = np.ascontiguousarray(spots, np.float32)
    n, box = sp.shape[0], sp.shape[1]
    th = np.empty((n, 6), np.float32)
    info = np.empty(n, np.int32)
    nfev = np.empty(n, np.int32)
    rc = lib.sim_lq(sp.ctypes.data, n, box, int(variant), th.ctypes.data, info.ctypes.data,
                    nfev.ctypes.data)
    assert rc == 0, rc
    return th, info, nfev


def stats(th, nfev, oth, onfev):
    d = th.astype(np.float64) - oth.astype(np.float64)
    rms = np.sqrt((d ** 2).mean(0))
    rel = np.sqrt(((d / np.maximum(np.abs(oth), 1e-6)) ** 2).mean(0))
    return {
        "same_nfev": float((nfev == onfev).mean()),
        "bit_identical_rows": float((th.view(np.uint32) == oth.view(np.uint32)).all(1).mean()),
        "rms_x_y_sx_sy_px": [float(rms[k]) for k in (0, 1, 4, 5)],
        "rel_rms_photons_bg": [float(rel[k]) for k in (2, 3)],
        "max_abs_x_y_sx_sy_px": [float(np.abs(d[:, k]).max()) for k in (0, 1, 4, 5)],
    }


if __name__ == "__main__":
    import oracle
    from picasso_b200 import testing

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    box = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    sp = testing.synthetic_spots(n, box, seed=77)
    oth, oinfo, onfev = oracle.fit_spots_lq(sp, nthreads=8, return_info=True)
    for variant in (0, 1):
        t0 = time.time()
        th, info, nfev = sim(sp, variant)
        print(variant, f"{time.time() - t0:.1f}s", stats(th, nfev, oth, onfev), flush=True)
