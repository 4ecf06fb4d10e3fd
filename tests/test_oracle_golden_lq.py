"""Pins oracle/lq_oracle.c (start values + float32-rounded model + MINPACK lmdif
restatement) to the real reference: golden thetas come from
picasso.gausslq.fit_spots, i.e. scipy.optimize.leastsq, run in the build container."""
import os

import numpy as np
import pytest


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "lq.npz"))


@pytest.mark.parametrize("box", [5, 7, 9, 11, 13])
def test_lq_oracle_poisson_spots(oracle, gold, box):
    spots = gold[f"b{box}_spots_u16"].astype(np.float32)
    th, info, nfev = oracle.fit_spots_lq(spots, nthreads=4, return_info=True)
    g = gold[f"b{box}_thetas"]
    assert (_bits(th) == _bits(g)).all(axis=1).mean() >= 0.999
    np.testing.assert_allclose(th, g, rtol=1e-4, atol=1e-4)
    assert set(np.unique(info)) <= {1, 2, 3, 4}
    assert 10 <= nfev.mean() <= 60


@pytest.mark.parametrize("key", ["float", "movie"])
def test_lq_oracle_float_spots(oracle, gold, key):
    th = oracle.fit_spots_lq(gold[f"{key}_spots"])
    g = gold[f"{key}_thetas"]
    assert (_bits(th) == _bits(g)).all(axis=1).mean() >= 0.999
    np.testing.assert_allclose(th, g, rtol=1e-4, atol=1e-4)


def test_lq_oracle_bundled_testdata(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "testdata.npz"))
    th = oracle.fit_spots_lq(g["spots"])
    assert (_bits(th) == _bits(g["lq_thetas"])).all()
    # SURVEY.md 8c known answer: mean theta of gausslq.fit_spots on the bundled movie
    np.testing.assert_allclose(th.mean(0), [-0.0310343, 0.1060921, 10957.64, 253.0364, 0.8695092,
                                            0.8668158], rtol=2e-6)
