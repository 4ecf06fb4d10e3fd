"""The per-spot arithmetic of the thread-per-spot CUDA MLE path
(picasso_b200/csrc/mle_tps_core.cuh), compiled for the host (tests/host_sim) and checked
against the golden vectors produced by the REAL reference (tools/gen_golden.py).

This pins the kernel's arithmetic -- separable edge evaluation, row-factorized Newton sums,
float32 or float64 per-pixel sums, Cholesky CRLB -- on the CPU-only build box: trajectory parity
(same iteration counts), the 1e-4 px RMS bar of BASELINE.json, CRLB and log-likelihood.  The GPU
tests then only have to show that the kernels drive these functions correctly.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def sim():
    import sim_mle_tps

    sim_mle_tps.build()
    return sim_mle_tps.sim


def _check(sim, spots, gold, prefix, method, f32, eps=0.001, max_it=100, min_same=0.995):
    th, cr, ll, it, st = sim(spots, eps, max_it, method, f32)
    gth, gcr = gold[f"{prefix}{method}_thetas"], gold[f"{prefix}{method}_crlbs"]
    gll, git = gold[f"{prefix}{method}_logliks"], gold[f"{prefix}{method}_iterations"]
    same = it == git
    assert same.mean() >= min_same, same.mean()
    d = th.astype(np.float64) - gth
    rms = np.sqrt((d ** 2).mean(0))
    assert rms[[0, 1, 4, 5]].max() <= 1e-4, rms
    rel = np.sqrt(((d / np.maximum(np.abs(gth), 1e-6)) ** 2).mean(0))
    assert rel[[2, 3]].max() <= 2e-4, rel
    # same trajectory -> same result to float32 rounding
    assert np.abs(d[same][:, [0, 1, 4, 5]]).max() <= 2e-5
    nz = gcr != 0
    assert ((cr == 0) == ~nz)[same].all()
    with np.errstate(divide="ignore", invalid="ignore"):
        crl = np.abs(cr - gcr) / np.abs(gcr)
    assert np.nanmax(crl[same][nz[same]]) <= 1e-4
    dll = np.abs(ll[same] - gll[same])
    assert (dll <= 1e-3 + 2e-6 * np.abs(gll[same])).all(), dll.max()
    return th, gth


@pytest.mark.parametrize("f32", [0, 1, 2])
@pytest.mark.parametrize("method", ["sigmaxy", "sigma"])
def test_sim_config1(sim, golden_dir, method, f32):
    g = np.load(os.path.join(golden_dir, "mle_config1.npz"))
    spots = g["spots_u16"].astype(np.float32)
    th, gth = _check(sim, spots, g, "", method, f32, min_same=0.999)
    if not f32:
        # float64 pixel sums: most rows are bit-identical to the reference
        bit = (th.view(np.uint32) == gth.view(np.uint32)).all(1).mean()
        assert bit >= 0.7, bit


@pytest.mark.parametrize("f32", [0, 1, 2])
@pytest.mark.parametrize("method", ["sigmaxy", "sigma"])
@pytest.mark.parametrize("box", [5, 9, 11, 13])
def test_sim_boxes(sim, golden_dir, box, method, f32):
    g = np.load(os.path.join(golden_dir, "mle_boxes.npz"))
    spots = g[f"b{box}_spots_u16"].astype(np.float32)
    _check(sim, spots, g, f"b{box}_", method, f32)


@pytest.mark.parametrize("f32", [0, 1, 2])
@pytest.mark.parametrize("method", ["sigmaxy", "sigma"])
@pytest.mark.parametrize("tag,eps,max_it", [("e3", 1e-3, 100), ("e6", 1e-6, 100),
                                            ("it3", 1e-3, 3), ("it0", 1e-3, 0)])
def test_sim_float_spots(sim, golden_dir, tag, eps, max_it, method, f32):
    g = np.load(os.path.join(golden_dir, "mle_float_spots.npz"))
    # eps 1e-6 is ~4 float32 ulps of theta: the stopping trip is noise-dominated (more so with
    # float32 pixel sums), but every spot has converged, so all thetas must agree
    min_same = (0.7 if f32 else 0.9) if tag == "e6" else 0.98
    th, cr, ll, it, st = sim(g["spots"], eps, max_it, method, f32)
    git, gth = g[f"{tag}_{method}_iterations"], g[f"{tag}_{method}_thetas"]
    same = it == git
    assert same.mean() >= min_same, same.mean()
    np.testing.assert_allclose(th[same][:, [0, 1, 4, 5]], gth[same][:, [0, 1, 4, 5]], atol=2e-5)
    np.testing.assert_allclose(th[same][:, [2, 3]], gth[same][:, [2, 3]], rtol=2e-4)
    if tag == "e6":
        np.testing.assert_allclose(th[:, [0, 1]], gth[:, [0, 1]], atol=2e-5)
        np.testing.assert_allclose(th[:, [4, 5]], gth[:, [4, 5]], atol=1e-4)   # sigma: not a stop criterion of "sigma"
    if max_it == 0:
        assert (it == 0).all()
        np.testing.assert_array_equal(th, gth)      # start values are bit-exact


def test_table_erf_accuracy(sim):
    """The 97-interval erf table of the float32-pixel kernels (erf_table.cuh, tools/gen_erf_table.py)
    against math.erf: max abs error of erf(z)/2 below 2e-11 on [-7, 7] (the fit needs ~1e-9, DESIGN.md
    5.1), odd symmetry, saturation beyond |z| = 6, NaN propagation."""
    import ctypes as C
    import math

    import sim_mle_tps

    lib = C.CDLL(sim_mle_tps.SIM_LIB)
    lib.sim_half_erf.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.uniform(-7, 7, 200_000), np.arange(-97, 98) / 16.0,
                        np.arange(-97, 98) / 16.0 + 1 / 32, [0.0, -0.0, 6.0, -6.0, 1e300, -1e300, np.inf, -np.inf]])
    out = np.empty_like(z)
    lib.sim_half_erf(z.ctypes.data, len(z), out.ctypes.data)
    ref = np.array([0.5 * math.erf(v) for v in z])
    assert np.abs(out - ref).max() <= 2e-11, np.abs(out - ref).max()
    np.testing.assert_array_equal(out[np.abs(z) >= 6], 0.5 * np.sign(z[np.abs(z) >= 6]))
    zn = np.array([np.nan, 0.3])
    on = np.empty(2)
    lib.sim_half_erf(zn.ctypes.data, 2, on.ctypes.data)
    assert np.isnan(on[0]) and abs(on[1] - 0.5 * math.erf(0.3)) <= 2e-11


def test_erf_table_is_the_generators_output():
    """csrc/erf_table.cuh is exactly what tools/gen_erf_table.py prints (no hand edits)."""
    import subprocess

    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_erf_table.py")], capture_output=True,
                         text=True, check=True).stdout
    with open(os.path.join(ROOT, "picasso_b200", "csrc", "erf_table.cuh")) as f:
        assert f.read() == out

