"""Compare the host build of the thread-per-spot MLE arithmetic (tests/host_sim) with the
oracle: iteration-count agreement, RMS, bit-identical rows.  No GPU needed.

    python tools/sim_mle_tps.py [n_spots]
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
SIM_LIB = os.path.join(SIM_DIR, "libmle_tps_sim.so")


def build():
    src = os.path.join(SIM_DIR, "mle_tps_sim.cpp")
    core = os.path.join(ROOT, "picasso_b200", "csrc", "mle_tps_core.cuh")
    if (not os.path.exists(SIM_LIB)
            or os.path.getmtime(SIM_LIB) < max(os.path.getmtime(src), os.path.getmtime(core))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC",
                               "-shared", "-x", "c++", src, "-o", SIM_LIB])
    lib = C.CDLL(SIM_LIB)
    vp = C.c_void_p
    lib.sim_mle_tps.argtypes = [vp, C.c_longlong, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                vp, vp, vp, vp, vp]
    return lib


def sim(spots, eps, max_it, method, f32):
    lib = build()
    n, box = spots.shape[0], spots.shape[1]
    th = np.empty((n, 6), np.float32)
    cr = np.empty((n, 6), np.float32)
    ll = np.empty(n, np.float32)
    it = np.empty(n, np.int32)
    st = np.empty(n, np.int32)
    sp = np.ascontiguousarray(spots, np.float32)
    rc = lib.sim_mle_tps(sp.ctypes.data, n, box, eps, max_it, 1 if method == "sigmaxy" else 0,
                         int(f32), th.ctypes.data, cr.ctypes.data, ll.ctypes.data, it.ctypes.data,
                         st.ctypes.data)
    assert rc == 0, rc
    return th, cr, ll, it, st


def report(spots, method, f32, eps=0.001, max_it=100):
    import oracle

    t0 = time.time()
    th, cr, ll, it, st = sim(spots, eps, max_it, method, f32)
    t1 = time.time()
    oth, ocr, oll, oit = oracle.gaussmle(spots, eps, max_it, method, nthreads=8)
    same = it == oit
    d = th.astype(np.float64) - oth
    rms = np.sqrt((d ** 2).mean(0))
    bit = (th.view(np.uint32) == oth.view(np.uint32)).all(1).mean()
    dsame = np.abs(d[same][:, [0, 1, 4, 5]]).max()
    nz = ocr != 0
    with np.errstate(divide="ignore", invalid="ignore"):
        crl = np.abs(cr - ocr) / np.abs(ocr)
    zeros_ok = ((cr == 0) == ~nz)[same].all()
    dll = np.abs(ll[same] - oll[same])
    llok = (dll <= 1e-3 + 2e-6 * np.abs(oll[same])).all()
    print(f"box {spots.shape[1]} {method} f32={f32}: it_match {same.mean():.5f} rms xy {rms[0]:.2e} "
          f"{rms[1]:.2e} s {rms[4]:.2e} {rms[5]:.2e} relN {np.sqrt(((d[:, 2] / oth[:, 2]) ** 2).mean()):.2e} "
          f"relbg {np.sqrt(((d[:, 3] / oth[:, 3]) ** 2).mean()):.2e} bit {bit:.3f} maxd_same {dsame:.2e} "
          f"crl_max {np.nanmax(crl[same][nz[same]]):.2e} zeros_ok {zeros_ok} ll_ok {llok} "
          f"dllmax {dll.max():.2e} ({t1 - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    from picasso_b200 import testing

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    sp = testing.synthetic_spots(n, 7, seed=3)
    for m in ("sigmaxy", "sigma"):
        for f32 in (0, 1, 2):
            report(sp, m, f32)
