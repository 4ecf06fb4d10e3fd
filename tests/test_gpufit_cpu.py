"""The Gpufit path on the CPU: the independent C restatement of Gpufit 1.2.0's LM algorithm
(oracle/gpufit_oracle.c; parity with the Windows binary is unpinned -- see its header) against
(a) the host build of the kernel's per-spot code (picasso_b200/csrc/gpufit_core.cuh), (b) ground truth
in the regime of the reference's own LQ tests, (c) the MINPACK path within the LQ tolerance."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "host_sim")


@pytest.fixture(scope="module")
def sim():
    src = os.path.join(SIM_DIR, "gpufit_sim.cpp")
    lib = os.path.join(SIM_DIR, "libgpufit_sim.so")
    core = os.path.join(ROOT, "picasso_b200", "csrc", "gpufit_core.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               src, "-o", lib])
    l = C.CDLL(lib)
    l.sim_gpufit.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 4

    def run(spots):
        sp = np.ascontiguousarray(spots, np.float32)
        n = len(sp)
        p = np.zeros((n, 6), np.float32); st = np.zeros(n, np.int32); chi = np.zeros(n, np.float32)
        nit = np.zeros(n, np.int32)
        assert l.sim_gpufit(sp.ctypes.data, n, sp.shape[1], 1e-2, 20, p.ctypes.data, st.ctypes.data,
                            chi.ctypes.data, nit.ctypes.data) == 0
        return p, st, chi, nit

    return run


@pytest.mark.parametrize("box", [5, 7, 9, 13])
def test_kernel_core_equals_independent_restatement(sim, oracle, box):
    from picasso_b200 import testing

    spots = testing.synthetic_spots(4000, box, seed=60 + box)
    p, st, chi, nit = sim(spots)
    op, ost, ochi, onit = oracle.fit_spots_gpufit(spots, nthreads=4, return_info=True)
    # same algorithm, same float32 operation order, no fused multiply-adds on either side
    np.testing.assert_array_equal(nit, onit)
    np.testing.assert_array_equal(st, ost)
    assert p.tobytes() == op.tobytes()
    assert chi.tobytes() == ochi.tobytes()


def test_oracle_ground_truth_and_lq_agreement(oracle):
    from picasso_b200 import testing

    grid = np.arange(-3, 4, dtype=np.float64)
    g1 = np.exp(-0.5 * grid ** 2) / np.sqrt(2 * np.pi)
    spot = (5000 * np.outer(g1, g1) + 10).astype(np.float32)
    ph, x, y, sx, sy, bg = oracle.fit_spots_gpufit(spot[None])[0]
    assert abs(x - 3) < 1e-3 and abs(y - 3) < 1e-3 and abs(sx - 1) < 1e-3 and abs(ph - 5000) / 5000 < 5e-3
    spots, truth = testing.synthetic_spots(5000, 7, seed=5, return_truth=True)
    p, st, chi, nit = oracle.fit_spots_gpufit(spots, nthreads=4, return_info=True)
    lq = oracle.fit_spots_lq(spots, nthreads=4)
    assert (st == 0).all() and 2 <= nit.mean() <= 6
    assert np.sqrt(np.mean((p[:, 1] - 3 - lq[:, 0]) ** 2)) < 5e-3
    # as accurate as the MINPACK path against the simulated positions
    e_gp = np.sqrt(np.mean((p[:, 1] - truth[:, 0]) ** 2))
    e_lq = np.sqrt(np.mean((lq[:, 0] + 3 - truth[:, 0]) ** 2))
    assert abs(e_gp - e_lq) < 2e-3 and e_gp < 0.06
