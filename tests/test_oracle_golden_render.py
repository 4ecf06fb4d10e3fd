"""Pins oracle/render_oracle.c to images rendered by the real reference
(picasso.render.render): same n, bit-identical float32 images (same libm, same
accumulation order)."""
import os

import numpy as np
import pytest

CASES = {
    "full_os8": dict(oversampling=8),
    "view_os5": dict(oversampling=5, viewport=((4.5, 3.25), (20.125, 30.75))),
    "os1_mbw1": dict(oversampling=1, min_blur_width=1),
    "os2p5_mbw": dict(oversampling=2.5, min_blur_width=0.1),
}
INFO = [{"Height": 24, "Width": 32, "Frames": 100, "Pixelsize": 130}]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "render.npz"))


@pytest.mark.parametrize("bm", [None, "gaussian", "gaussian_iso"])
@pytest.mark.parametrize("tag", list(CASES))
def test_render_oracle_matches_reference(oracle, gold, tag, bm):
    locs = {k: gold[k] for k in ("x", "y", "lpx", "lpy")}
    n, img = oracle.render(locs, INFO, blur_method=bm, **CASES[tag])
    assert n == int(gold[f"{tag}_{bm}_n"])
    g = gold[f"{tag}_{bm}_image"]
    assert img.shape == g.shape and img.dtype == np.float32
    same = (img.view(np.uint32) == g.view(np.uint32)).mean()
    assert same >= 0.9999, same
    np.testing.assert_allclose(img, g, rtol=1e-6, atol=1e-9)
    if bm is None:
        assert img.sum() == n


def test_render_oracle_bundled_testdata(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "testdata.npz"))
    locs = {"x": g["locs_x"], "y": g["locs_y"], "lpx": g["locs_lpx"], "lpy": g["locs_lpy"]}
    info = [{"Height": 32, "Width": 32, "Frames": 100, "Pixelsize": 130}]
    for bm, total in ((None, 30.0), ("gaussian", 41.785927), ("gaussian_iso", 41.842075)):
        n, img = oracle.render(locs, info, oversampling=20, blur_method=bm)
        assert n == 30 and img.shape == (640, 640)
        np.testing.assert_array_equal(img, g[f"render_{bm}_image"])
        assert abs(float(img.sum()) - total) < 1e-3      # SURVEY.md 8c known answers


def test_render_oracle_errors(oracle):
    with pytest.raises(Exception, match="blur_method not understood."):
        oracle.render({"x": np.zeros(1), "y": np.zeros(1)}, [{"Height": 4, "Width": 4}],
                      blur_method="nope")
