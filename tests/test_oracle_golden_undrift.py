"""Pins the numpy undrift oracle (oracle/undrift_oracle.py + render oracle) to the real
reference on a reduced instance (8 segments of 64x72, 28 pairs)."""
import os

import numpy as np
import pytest

from oracle import undrift_oracle as uo


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "undrift.npz"))


@pytest.fixture(scope="module")
def problem(gold):
    locs = {k: gold[k] for k in ("frame", "x", "y", "lpx", "lpy")}
    H, W, F = (int(v) for v in gold["info_hwf"])
    return locs, [{"Height": H, "Width": W, "Frames": F, "Pixelsize": 130}]


def test_segment_matches_reference(oracle, gold, problem):
    locs, info = problem
    bounds, segs = uo.segment(locs, info, 100, oracle.render,
                              {"blur_method": "gaussian", "min_blur_width": 1})
    assert bounds.dtype == np.uint32
    np.testing.assert_array_equal(bounds, gold["bounds"])
    assert segs.dtype == np.float64 and segs.shape == gold["segments"].shape
    np.testing.assert_allclose(segs, gold["segments"], rtol=1e-6, atol=1e-9)


def test_xcorr_and_pair_shifts_match_reference(gold):
    segs = gold["segments"].astype(np.float64)
    np.testing.assert_allclose(uo.xcorr(segs[0], segs[1]), gold["xcorr_0_1"], rtol=1e-12, atol=1e-12)
    sy, sx = uo.pair_shifts(segs, 32)
    np.testing.assert_allclose(sy, gold["pair_shift_y"], atol=1e-9)
    np.testing.assert_allclose(sx, gold["pair_shift_x"], atol=1e-9)
    np.testing.assert_allclose(uo.get_image_shift(segs[0], segs[3], 5, None), gold["shift_noroi_0_3"],
                               atol=1e-9)
    ry, rx = uo.rcc(segs, 32)
    np.testing.assert_allclose(ry, gold["rcc_shift_y"], atol=1e-9)
    np.testing.assert_allclose(rx, gold["rcc_shift_x"], atol=1e-9)


def test_undrift_matches_reference(oracle, gold, problem):
    locs, info = problem
    drift, xn, yn = uo.undrift(locs, info, 100, oracle.render)
    np.testing.assert_allclose(drift[:, 0], gold["drift_x"], atol=1e-7)
    np.testing.assert_allclose(drift[:, 1], gold["drift_y"], atol=1e-7)
    np.testing.assert_allclose(xn, gold["undrifted_x"], atol=1e-5)
    np.testing.assert_allclose(yn, gold["undrifted_y"], atol=1e-5)


def test_zero_image_and_edge_window():
    z = np.zeros((16, 16))
    a = np.zeros((16, 16)); a[8, 8] = 1
    assert uo.get_image_shift(z, a, 5, None) == (0, 0)
    # minimize_shifts: exact recovery (reference tests/test_lib.py:511-551)
    true = np.array([0.0, 0.5, -0.25, 1.0])
    sx = np.zeros((4, 4)); sy = np.zeros((4, 4))
    for i in range(3):
        for j in range(i + 1, 4):
            sx[i, j] = true[j] - true[i]
            sy[i, j] = 2 * (true[j] - true[i])
    y, x = uo.minimize_shifts(sx, sy)
    np.testing.assert_allclose(x, true, atol=1e-9)
    np.testing.assert_allclose(y, 2 * true, atol=1e-9)


def test_config5_reduced_instance_matches_reference(oracle, golden_dir):
    """Reduced BASELINE config 5 (20 segments x 1024^2, tools/gen_golden.py undrift_c5): the
    oracle's segment renders and a sample of its per-pair shifts against the real reference."""
    g = np.load(os.path.join(golden_dir, "undrift_c5.npz"))
    locs = {k: g[k] for k in ("frame", "x", "y", "lpx", "lpy")}
    H, W, F = (int(v) for v in g["info_hwf"])
    info = [{"Height": H, "Width": W, "Frames": F, "Pixelsize": 130}]
    bounds, segs = uo.segment(locs, info, 100, oracle.render,
                              {"blur_method": "gaussian", "min_blur_width": 1})
    np.testing.assert_array_equal(bounds, g["bounds"])
    np.testing.assert_allclose(segs.sum((1, 2)), g["segment_sums"], rtol=1e-9)
    for (i, j) in ((0, 1), (0, 19), (7, 12), (18, 19)):
        sy, sx = uo.get_image_shift(segs[i], segs[j], 5, 32)
        assert abs(sy - g["pair_shift_y"][i, j]) < 1e-9 and abs(sx - g["pair_shift_x"][i, j]) < 1e-9
