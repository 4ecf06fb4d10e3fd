"""Pins oracle/localize_oracle.c to the real reference (localize.py) on the golden
vectors from tools/gen_golden.py identify / testdata: detections must be
bit-exact (coordinates, order, float32 net gradient)."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "identify.npz"))


@pytest.mark.parametrize("box", [3, 5, 7, 9, 11, 13])
def test_local_maxima_tie_rule(oracle, gold, box):
    y, x = oracle.local_maxima(gold[f"ties_b{box}_frame"], box)
    np.testing.assert_array_equal(y, gold[f"ties_b{box}_y"])
    np.testing.assert_array_equal(x, gold[f"ties_b{box}_x"])
    assert len(y) > 0


@pytest.mark.parametrize("box", [5, 7, 9])
def test_net_gradient_wraparound(oracle, gold, box):
    y, x, ng = oracle.identify_in_image(gold[f"wrap_b{box}_frame"], -1e30, box)
    np.testing.assert_array_equal(y, gold[f"wrap_b{box}_y"])
    np.testing.assert_array_equal(x, gold[f"wrap_b{box}_x"])
    assert (ng.view(np.uint32) == gold[f"wrap_b{box}_ng"].view(np.uint32)).all()


@pytest.mark.parametrize("box,mng", [(7, 5000), (9, 8000), (5, 3000)])
def test_identify_movie_and_get_spots(oracle, gold, box, mng):
    fr, x, y, ng = oracle.identify_movie(gold["movie"], mng, box)
    tag = f"mov_b{box}"
    np.testing.assert_array_equal(fr, gold[f"{tag}_frame"])
    np.testing.assert_array_equal(x, gold[f"{tag}_x"])
    np.testing.assert_array_equal(y, gold[f"{tag}_y"])
    assert (ng.view(np.uint32) == gold[f"{tag}_ng"].view(np.uint32)).all()
    cam = {"Baseline": 100, "Sensitivity": 0.45, "Gain": 2}
    spots = oracle.get_spots(gold["movie"], fr, x, y, box, cam)
    assert (spots.view(np.uint32) == gold[f"{tag}_spots"].view(np.uint32)).all()


def test_identify_roi_and_frame_bounds(oracle, gold):
    roi = tuple(map(tuple, gold["roi"]))
    fb = tuple(gold["roi_frame_bounds"])
    fr, x, y, ng = oracle.identify_movie(gold["movie"], 5000, 7, roi, fb)
    np.testing.assert_array_equal(fr, gold["roi_frame"])
    np.testing.assert_array_equal(x, gold["roi_x"])
    np.testing.assert_array_equal(y, gold["roi_y"])
    assert (ng.view(np.uint32) == gold["roi_ng"].view(np.uint32)).all()
    assert fr.min() >= 3 and fr.max() <= 8


def test_bundled_testdata_known_answers(oracle, golden_dir):
    """SURVEY.md 8c known answers on the reference's tests/data/testdata.raw."""
    g = np.load(os.path.join(golden_dir, "testdata.npz"))
    fr, x, y, ng = oracle.identify_movie(g["movie"], 5000, 7)
    assert len(fr) == 30 and fr.sum() == 1790 and x.sum() == 426 and y.sum() == 481
    assert abs(float(ng.sum()) - 739791.25) < 0.5
    np.testing.assert_array_equal(fr, g["ids_frame"])
    assert (ng.view(np.uint32) == g["ids_ng"].view(np.uint32)).all()
    spots = oracle.get_spots(g["movie"], fr, x, y, 7, {"Baseline": 0, "Sensitivity": 1, "Gain": 1})
    np.testing.assert_array_equal(spots, g["spots"])
    assert spots.sum() == 700636
    th, cr, ll, it = oracle.gaussmle(spots, 0.001, 100, "sigmaxy")
    assert (th.view(np.uint32) == g["mle_thetas"].view(np.uint32)).all()
    np.testing.assert_array_equal(it, g["mle_iterations"])
    np.testing.assert_allclose(cr, g["mle_crlbs"], rtol=5e-6)
    th, cr, ll, it = oracle.gaussmle(spots, 0.001, 100, "sigma")
    assert (th.view(np.uint32) == g["mles_thetas"].view(np.uint32)).all()
