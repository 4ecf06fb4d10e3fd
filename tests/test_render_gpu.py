"""GPU parity: CUDA renderer (through the C ABI) vs images from the real reference
(golden) and vs the CPU oracle.  `n` and the no-blur histogram are exact; blurred
images agree to float32 summation-order rounding (per-pixel rtol 1e-4 above 1e-3 of
the maximum)."""
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import render as pbrender

pytestmark = pytest.mark.gpu

CASES = {
    "full_os8": dict(oversampling=8),
    "view_os5": dict(oversampling=5, viewport=((4.5, 3.25), (20.125, 30.75))),
    "os1_mbw1": dict(oversampling=1, min_blur_width=1),
    "os2p5_mbw": dict(oversampling=2.5, min_blur_width=0.1),
}
INFO = [{"Height": 24, "Width": 32, "Frames": 100, "Pixelsize": 130}]


def _close(img, ref):
    assert img.shape == ref.shape and img.dtype == np.float32
    big = ref > 1e-3 * ref.max()
    np.testing.assert_allclose(img[big], ref[big], rtol=1e-4)
    np.testing.assert_allclose(img, ref, rtol=1e-4, atol=1e-6 * max(1.0, float(ref.max())))
    np.testing.assert_allclose(img.sum(dtype=np.float64), ref.sum(dtype=np.float64), rtol=1e-5)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "render.npz"))


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("bm", [None, "gaussian", "gaussian_iso"])
@pytest.mark.parametrize("tag", list(CASES))
def test_render_golden(gold, tag, bm):
    locs = pd.DataFrame({k: gold[k] for k in ("x", "y", "lpx", "lpy")})
    n, img = pbrender.render(locs, INFO, blur_method=bm, **CASES[tag])
    assert n == int(gold[f"{tag}_{bm}_n"])
    ref = gold[f"{tag}_{bm}_image"]
    if bm is None:
        np.testing.assert_array_equal(img, ref)
        assert img.sum() == n
    else:
        _close(img, ref)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_render_bundled_testdata(golden_dir):
    g = np.load(os.path.join(golden_dir, "testdata.npz"))
    locs = pd.DataFrame({"x": g["locs_x"], "y": g["locs_y"], "lpx": g["locs_lpx"],
                         "lpy": g["locs_lpy"]})
    info = [{"Height": 32, "Width": 32, "Frames": 100, "Pixelsize": 130}]
    for bm in (None, "gaussian", "gaussian_iso"):
        n, img = pbrender.render(locs, info, oversampling=20, blur_method=bm)
        assert n == 30 and img.shape == (640, 640)
        _close(img, g[f"render_{bm}_image"])


@pytest.fixture
def render_impl():
    """Select the accumulation pass of the binned path for one test (0 = tile atomics, the default;
    1 = warp-owned strips) and restore the default afterwards."""
    from picasso_b200 import _lib

    lib = _lib.load()
    default = lib.pb_render_get_impl()

    def select(impl):
        _lib.check(lib.pb_render_set_impl(impl))

    yield select
    lib.pb_render_set_impl(default)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("bm", ["gaussian", "gaussian_iso", None])
def test_render_large_tiled_vs_oracle(oracle, bm, impl, render_impl):
    """300k localisations at oversampling 20 -> the binned shared-memory paths (both accumulation
    passes: tile atomics and warp-owned strips)."""
    if bm is None and impl == 1:
        pytest.skip("the histogram has no binned path")
    render_impl(impl)
    rng = np.random.default_rng(2)
    n = 300_000
    locs = pd.DataFrame({"x": rng.uniform(0, 64, n).astype(np.float32),
                         "y": rng.uniform(0, 48, n).astype(np.float32),
                         "lpx": rng.uniform(0.02, 0.08, n).astype(np.float32),
                         "lpy": rng.uniform(0.02, 0.08, n).astype(np.float32)})
    info = [{"Height": 48, "Width": 64, "Frames": 1, "Pixelsize": 130}]
    k, img = pbrender.render(locs, info, oversampling=20, blur_method=bm)
    ok, oimg = oracle.render(locs, info, oversampling=20, blur_method=bm)
    assert k == ok and img.shape == (960, 1280)
    if bm is None:
        np.testing.assert_array_equal(img, oimg)
    else:
        _close(img, oimg)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("impl", [0, 1])
def test_render_binned_wide_windows(oracle, impl, render_impl):
    """Windows wider than the per-thread column cache (16 px) / taller than two strips among ordinary
    ones: both accumulation passes route them through their generic branches."""
    render_impl(impl)
    rng = np.random.default_rng(5)
    n = 200_000
    lp = rng.uniform(0.02, 0.08, (2, n)).astype(np.float32)
    lp[:, :2000] = rng.uniform(0.15, 0.45, (2, 2000)).astype(np.float32)
    locs = pd.DataFrame({"x": rng.uniform(0, 64, n).astype(np.float32),
                         "y": rng.uniform(0, 48, n).astype(np.float32), "lpx": lp[0], "lpy": lp[1]})
    info = [{"Height": 48, "Width": 64, "Frames": 1, "Pixelsize": 130}]
    for bm in ("gaussian", "gaussian_iso"):
        k, img = pbrender.render(locs, info, oversampling=20, blur_method=bm)
        ok, oimg = oracle.render(locs, info, oversampling=20, blur_method=bm)
        assert k == ok
        _close(img, oimg)


def test_render_api_contract():
    locs = pd.DataFrame({"x": np.float32([1.5, 2.5]), "y": np.float32([1.5, 2.5]),
                         "lpx": np.float32([0.1, 0.1]), "lpy": np.float32([0.1, 0.1])})
    info = [{"Height": 8, "Width": 8, "Frames": 1, "Pixelsize": 100}]
    with pytest.warns(DeprecationWarning):
        n, img = pbrender.render(locs, info, oversampling=4)
    assert n == 2 and img.shape == (32, 32) and img.sum() == 2
    # oversampling=k == disp_px_size=pixelsize/k (reference test_render.py:133-145)
    n2, img2 = pbrender.render(locs, info, disp_px_size=25)
    np.testing.assert_array_equal(img, img2)
    with pytest.raises(Exception, match="blur_method not understood."):
        pbrender.render(locs, info, disp_px_size=25, blur_method="nope")
    with pytest.raises(KeyError):
        pbrender.render(locs, [{"Height": 8, "Width": 8}], disp_px_size=25)
    with pytest.raises(KeyError):      # dict info without a viewport: info[0] fails like the reference
        pbrender.render(locs, {"Pixelsize": 100}, disp_px_size=25)
    # empty locs
    n0, img0 = pbrender.render(locs.iloc[:0], info, disp_px_size=25, blur_method="gaussian")
    assert n0 == 0 and img0.shape == (32, 32) and img0.sum() == 0
    # strict viewport: a loc exactly on the border is out
    edge = pd.DataFrame({"x": np.float32([0.0, 8.0, 4.0]), "y": np.float32([4.0, 4.0, 4.0]),
                         "lpx": np.float32([0.1] * 3), "lpy": np.float32([0.1] * 3)})
    n3, _ = pbrender.render(edge, info, disp_px_size=25)
    assert n3 == 1


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_render_config4_subsample_vs_oracle(oracle):
    """BASELINE config 4 geometry (512 x 512 camera pixels at oversampling 20 -> 10240 x 10240,
    x, y ~ U(0, 512), lp ~ U(0.02, 0.08), seed 2) on a 2 M-localization sub-sample
    against the oracle: exact n, rtol 1e-4 on pixels above 1e-3 of the maximum."""
    rng = np.random.default_rng(2)
    n = 2_000_000
    locs = pd.DataFrame({"x": rng.uniform(0, 512, n).astype(np.float32),
                         "y": rng.uniform(0, 512, n).astype(np.float32),
                         "lpx": rng.uniform(0.02, 0.08, n).astype(np.float32),
                         "lpy": rng.uniform(0.02, 0.08, n).astype(np.float32)})
    info = [{"Height": 512, "Width": 512, "Frames": 1, "Pixelsize": 130}]
    k, img = pbrender.render(locs, info, oversampling=20, blur_method="gaussian")
    ok, oimg = oracle.render(locs, info, oversampling=20, blur_method="gaussian")
    assert k == ok and img.shape == (10240, 10240) and img.dtype == np.float32
    big = oimg > 1e-3 * oimg.max()
    np.testing.assert_allclose(img[big], oimg[big], rtol=1e-4)
    assert abs(float(img.sum(dtype=np.float64)) - float(oimg.sum(dtype=np.float64))) <= 1e-5 * float(oimg.sum(dtype=np.float64))
    kh, hist = pbrender.render(locs, info, oversampling=20, blur_method=None)
    okh, ohist = oracle.render(locs, info, oversampling=20, blur_method=None)
    assert kh == okh
    np.testing.assert_array_equal(hist, ohist)
