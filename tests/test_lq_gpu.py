"""GPU parity: CUDA least-squares fit (MINPACK-faithful lmdif per thread) vs golden
vectors from the real reference (scipy.optimize.leastsq) and vs the CPU oracle.

Stated LQ tolerance: the optimiser stops after ~3 LM iterations at ftol=xtol=1e-2,
so parity means following the same trajectory: >= 99 % of spots with the same number
of residual evaluations (nfev) and, on those, x/y/sigma within 1e-4 px, photons/bg
within 1e-4 relative."""
import os

import numpy as np
import pytest

from picasso_b200 import gausslq, testing

pytestmark = pytest.mark.gpu


def _check(th, ref, same_frac_min=0.99):
    d = np.abs(th.astype(np.float64) - ref)
    tol = np.array([1e-4, 1e-4, 0, 0, 1e-4, 1e-4]) + 1e-4 * np.abs(ref) * np.array([0, 0, 1, 1, 0, 0])
    ok = (d <= tol).all(1)
    assert ok.mean() >= same_frac_min, ok.mean()
    rms = np.sqrt((d[:, [0, 1, 4, 5]] ** 2).mean(0))
    assert rms.max() <= 2e-3, rms          # the few off-trajectory spots stay within LQ's own tolerance


@pytest.mark.parametrize("box", [5, 7, 9, 11, 13])
def test_lq_matches_reference_golden(golden_dir, box):
    g = np.load(os.path.join(golden_dir, "lq.npz"))
    spots = g[f"b{box}_spots_u16"].astype(np.float32)
    th = gausslq.fit_spots(spots)
    assert th.shape == (len(spots), 6) and th.dtype == np.float32
    _check(th, g[f"b{box}_thetas"])


@pytest.mark.parametrize("key", ["float", "movie"])
def test_lq_float_spots_golden(golden_dir, key):
    g = np.load(os.path.join(golden_dir, "lq.npz"))
    _check(gausslq.fit_spots(g[f"{key}_spots"]), g[f"{key}_thetas"], 0.97)


def test_lq_vs_oracle_trajectory(oracle):
    spots = testing.synthetic_spots(20000, 7, seed=77)
    th, info, nfev = gausslq._fit(spots, want_info=True)
    oth, oinfo, onfev = oracle.fit_spots_lq(spots, nthreads=8, return_info=True)
    assert (nfev == onfev).mean() >= 0.99
    same = nfev == onfev
    np.testing.assert_allclose(th[same], oth[same], rtol=2e-4, atol=2e-4)
    assert set(np.unique(info)) <= {1, 2, 3, 4}


def test_lq_api_shapes_and_gpufit_layout():
    spots = testing.synthetic_spots(257, 7, seed=5)     # ragged last CTA
    th = gausslq.fit_spots(spots)
    one = gausslq.fit_spot(spots[3])
    np.testing.assert_allclose(one, th[3], rtol=1e-6)
    fs = gausslq.fit_spots_parallel(spots, asynch=True)
    np.testing.assert_array_equal(gausslq.fits_from_futures(fs), th)
    np.testing.assert_array_equal(gausslq.fit_spots_parallel(spots), th)
    gp = gausslq.fit_spots_gpufit(spots)
    np.testing.assert_allclose(gp[:, 0], th[:, 2])
    np.testing.assert_allclose(gp[:, 1], th[:, 0] + 3)
    np.testing.assert_allclose(gp[:, 5], th[:, 3])
    seen = []
    gausslq.fit_spots(spots[:10], seen.append)
    assert seen == list(range(10))
    assert gausslq.fit_spots(np.zeros((0, 7, 7), np.float32)).shape == (0, 6)


def test_lq_centered_spot_ground_truth():
    """Reference test_gausslq.py:38-50 regime: noiseless centred spot."""
    half = 3
    grid = np.arange(-half, half + 1, dtype=np.float64)
    g1 = np.exp(-0.5 * (grid / 1.0) ** 2) / np.sqrt(2 * np.pi)
    spot = (5000 * np.outer(g1, g1) + 10).astype(np.float32)
    x, y, ph, bg, sx, sy = gausslq.fit_spot(spot)
    assert abs(x) < 1e-3 and abs(y) < 1e-3
    assert abs(sx - 1) < 1e-3 and abs(sy - 1) < 1e-3
    assert abs(ph - 5000) / 5000 < 5e-3
