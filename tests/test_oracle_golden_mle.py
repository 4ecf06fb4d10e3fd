"""Pins the CPU oracle (oracle/mle_oracle.c) to the REAL reference: golden vectors
in tests/golden/ were produced by picasso.gaussmle.gaussmle imported from the
reference (tools/gen_golden.py).  thetas / iterations / log-likelihoods must be
bit-identical on the build container (same glibc libm); CRLBs agree to the
LAPACK-pinv vs Jacobi rounding (1 f32 ulp)."""
import os

import numpy as np
import pytest


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _check(orc, spots, gold, prefix, method, eps=0.001, max_it=100):
    th, cr, ll, it = orc.gaussmle(spots, eps, max_it, method, nthreads=4)
    gth, gcr = gold[f"{prefix}{method}_thetas"], gold[f"{prefix}{method}_crlbs"]
    gll, git = gold[f"{prefix}{method}_logliks"], gold[f"{prefix}{method}_iterations"]
    assert (it == git).mean() >= 0.999
    # bit-identical apart from (at most) libm last-bit differences on another host
    assert (_bits(th) == _bits(gth)).all(axis=1).mean() >= 0.999
    np.testing.assert_allclose(th, gth, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(ll, gll, rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(cr, gcr, rtol=5e-6)


@pytest.mark.parametrize("method", ["sigmaxy", "sigma"])
def test_oracle_config1(oracle, golden_dir, method):
    g = np.load(os.path.join(golden_dir, "mle_config1.npz"))
    spots = g["spots_u16"].astype(np.float32)
    _check(oracle, spots, g, "", method)


@pytest.mark.parametrize("method", ["sigmaxy", "sigma"])
@pytest.mark.parametrize("box", [5, 9, 11, 13, 15])
def test_oracle_boxes(oracle, golden_dir, box, method):
    g = np.load(os.path.join(golden_dir, "mle_boxes.npz"))
    spots = g[f"b{box}_spots_u16"].astype(np.float32)
    _check(oracle, spots, g, f"b{box}_", method)


@pytest.mark.parametrize("method", ["sigmaxy", "sigma"])
@pytest.mark.parametrize("tag,eps,max_it", [("e3", 1e-3, 100), ("e6", 1e-6, 100),
                                            ("it3", 1e-3, 3), ("it0", 1e-3, 0)])
def test_oracle_float_spots(oracle, golden_dir, tag, eps, max_it, method):
    g = np.load(os.path.join(golden_dir, "mle_float_spots.npz"))
    _check(oracle, g["spots"], g, f"{tag}_", method, eps, max_it)


def test_oracle_bad_method(oracle):
    with pytest.raises(ValueError, match="Method not available."):
        oracle.gaussmle(np.zeros((1, 7, 7), np.float32), 1e-3, 10, "bogus")
