"""GPU parity: cuFFT RCC + CUDA segment rendering vs the real reference (golden) and the
numpy oracle.  Tolerance: per-pair shifts, segment shifts and the per-frame drift within
1e-3 px (the reference's own tests allow 0.5 px)."""
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import imageprocess, lib, postprocess, testing

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore::DeprecationWarning")]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "undrift.npz"))


@pytest.fixture(scope="module")
def problem(gold):
    locs = pd.DataFrame({k: gold[k] for k in ("frame", "x", "y", "lpx", "lpy")})
    H, W, F = (int(v) for v in gold["info_hwf"])
    return locs, [{"Height": H, "Width": W, "Frames": F, "Pixelsize": 130}]


def test_n_segments_bankers_rounding():
    assert postprocess.n_segments([{"Frames": 250}], 100) == 2      # 2.5 -> 2
    assert postprocess.n_segments([{"Frames": 350}], 100) == 4      # 3.5 -> 4
    assert postprocess.n_segments([{"Frames": 1000}], 100) == 10


def test_segment_golden(gold, problem):
    locs, info = problem
    seen = []
    bounds, segs = postprocess.segment(locs, info, 100, {"blur_method": "gaussian",
                                                         "min_blur_width": 1}, seen.append)
    assert seen == list(range(9))
    assert bounds.dtype == np.uint32 and segs.dtype == np.float64
    np.testing.assert_array_equal(bounds, gold["bounds"])
    ref = gold["segments"]
    np.testing.assert_allclose(segs, ref, rtol=1e-4, atol=1e-6 * ref.max())
    # mass == locs minus the dropped last frame (reference test_undrift.py:305-312)
    kept = (locs["frame"] < bounds[-1]).sum()
    assert kept < len(locs)


def test_xcorr_and_shifts_golden(gold):
    segs = gold["segments"].astype(np.float64)
    xc = imageprocess.xcorr(segs[0], segs[1])
    ref = gold["xcorr_0_1"]
    assert xc.shape == ref.shape
    np.testing.assert_allclose(xc, ref, rtol=0, atol=2e-5 * np.abs(ref).max())
    assert np.unravel_index(xc.argmax(), xc.shape) == np.unravel_index(ref.argmax(), ref.shape)
    for (i, j) in ((0, 1), (0, 7), (3, 4)):
        sy, sx = imageprocess.get_image_shift(segs[i], segs[j], 5, 32)
        assert abs(sy - gold["pair_shift_y"][i, j]) < 1e-3 and abs(sx - gold["pair_shift_x"][i, j]) < 1e-3
    sy, sx = imageprocess.get_image_shift(segs[0], segs[3], 5, None)
    np.testing.assert_allclose([sy, sx], gold["shift_noroi_0_3"], atol=1e-3)
    assert imageprocess.get_image_shift(np.zeros((16, 16)), segs[0][:16, :16], 5) == (0, 0)


def test_rcc_golden(gold):
    segs = gold["segments"].astype(np.float64)
    seen = []
    shift_y, shift_x = imageprocess.rcc(segs, 32, seen.append)
    assert seen == list(range(29))
    np.testing.assert_allclose(shift_y, gold["rcc_shift_y"], atol=1e-3)
    np.testing.assert_allclose(shift_x, gold["rcc_shift_x"], atol=1e-3)


def test_undrift_golden_and_ground_truth(gold, problem):
    locs, info = problem
    drift, und = postprocess.undrift(locs, info, 100, display=False,
                                     segmentation_callback=lambda i: None,
                                     rcc_callback=lambda i: None)
    assert list(drift.columns) == ["x", "y"] and len(drift) == info[0]["Frames"]
    np.testing.assert_allclose(drift["x"].to_numpy(), gold["drift_x"], atol=1e-3)
    np.testing.assert_allclose(drift["y"].to_numpy(), gold["drift_y"], atol=1e-3)
    np.testing.assert_allclose(und["x"].to_numpy(), gold["undrifted_x"], atol=1e-3)
    np.testing.assert_allclose(und["y"].to_numpy(), gold["undrifted_y"], atol=1e-3)
    assert und is not locs and not np.array_equal(und["x"].to_numpy(), locs["x"].to_numpy())


def test_undrift_recovers_injected_drift():
    """Reference test_undrift.py:333-340: RCC recovers the injected drift within 0.5 px
    after de-meaning."""
    locs, info, truth = testing.synthetic_drift_locs(2000, 96, 96, n_clusters=60,
                                                     locs_per_frame=8, seed=5)
    drift, _ = postprocess.undrift(locs, info, 200, display=False,
                                   segmentation_callback=lambda i: None,
                                   rcc_callback=lambda i: None)
    for k, col in enumerate(("x", "y")):
        est = drift[col].to_numpy() - drift[col].mean()
        tru = truth[:, k] - truth[:, k].mean()
        assert np.abs(est - tru).max() < 0.5


def test_apply_drift_and_minimize_shifts():
    locs = pd.DataFrame({"frame": np.uint32([0, 1, 2]), "x": np.float32([1, 1, 1]),
                         "y": np.float32([2, 2, 2])})
    info = [{"Frames": 3}]
    out = postprocess.apply_drift(locs.copy(), info, drift=np.array([[0.1, 0.2], [0.2, 0.4], [0.3, 0.6]]))
    np.testing.assert_allclose(out["x"], [0.9, 0.8, 0.7], atol=1e-6)
    np.testing.assert_allclose(out["y"], [1.8, 1.6, 1.4], atol=1e-6)
    with pytest.raises(ValueError):
        postprocess.apply_drift(locs.copy(), info, drift=np.zeros((2, 2)))
    true = np.array([0.0, 0.5, -0.25, 1.0])
    sx = np.zeros((4, 4)); sy = np.zeros((4, 4))
    for i in range(3):
        for j in range(i + 1, 4):
            sx[i, j] = true[j] - true[i]
    y, x = lib.minimize_shifts(sx, sy)
    np.testing.assert_allclose(x, true, atol=1e-9)


def test_fused_undrift_equals_segment_plus_rcc(problem):
    """postprocess.undrift renders the segments on the device; the public two-step path
    (segment() -> rcc()) must give the same shifts."""
    locs, info = problem
    bounds, segs = postprocess.segment(locs, info, 100, {"blur_method": "gaussian",
                                                         "min_blur_width": 1}, lambda i: None)
    sy, sx = imageprocess.rcc(segs, 32, lambda i: None)
    fy, fx = imageprocess._rcc_of_locs(locs, info, bounds, 1, 32, lambda i: None)
    np.testing.assert_allclose(fy, sy, atol=2e-4)
    np.testing.assert_allclose(fx, sx, atol=2e-4)
    seg_cb, rcc_cb = [], []
    postprocess.undrift(locs, info, 100, display=False, segmentation_callback=seg_cb.append,
                        rcc_callback=rcc_cb.append)
    assert seg_cb == list(range(9)) and rcc_cb == list(range(29))


@pytest.mark.parametrize("shape,win", [((64, 64), (0, 0, 64, 64)), ((96, 80), (32, 24, 32, 32)),
                                       ((128, 66), (40, 10, 37, 21)), ((50, 36), (9, 2, 32, 32)),
                                       ((256, 256), (112, 112, 32, 32))])
def test_pruned_inverse_equals_cufft_and_numpy(shape, win):
    """The pruned inverse DFT of the window rows / columns (csrc/rcc.cu) and the cuFFT C2R + crop
    path give the same correlation window as numpy's float64 xcorr (imageprocess.py:27-50)."""
    import ctypes as C

    from picasso_b200 import _lib
    l = _lib.load()
    l.pb_rcc_set_mode.argtypes = [C.c_int]
    Y, X = shape
    Y0, X0, H, W = win
    rng = np.random.default_rng(5)
    segs = rng.poisson(2.0, (3, Y, X)).astype(np.float32)
    segs[1] = np.roll(segs[0], (2, -3), (0, 1)) + rng.poisson(0.3, (Y, X))
    out = {}
    try:
        for mode in (0, 1):
            assert l.pb_rcc_set_mode(mode) == 0
            w = np.zeros((3, H, W), np.float32)
            sums = np.zeros(3, np.float64)
            _lib.check(l.pb_rcc_windows(3, Y, X, _lib.ptr(segs), Y0, X0, H, W, _lib.ptr(w), _lib.ptr(sums)))
            out[mode] = w
    finally:
        l.pb_rcc_set_mode(-1)
    pairs = [(0, 1), (0, 2), (1, 2)]
    for k, (i, j) in enumerate(pairs):
        a, b = segs[i].astype(np.float64), segs[j].astype(np.float64)
        ref = np.fft.fftshift(np.real(np.fft.ifft2(np.fft.fft2(a) * np.conj(np.fft.fft2(b))))) / np.sqrt(a.size)
        ref = ref[Y0:Y0 + H, X0:X0 + W]
        tol = 5e-6 * np.abs(ref).max()
        np.testing.assert_allclose(out[0][k], ref, rtol=0, atol=tol)
        np.testing.assert_allclose(out[1][k], ref, rtol=0, atol=tol)
    assert l.pb_rcc_set_mode(7) != 0


def test_device_peak_fits_equal_host_fits():
    """pb_undrift_peaks_pairs (arg-max, 5x5 cut-out and Levenberg-Marquardt fit on the GPU) gives
    the shifts of the window path with host fits; records carry the 5x5 window for re-fits."""
    import ctypes as C

    from picasso_b200 import _lib
    locs, info, truth = testing.synthetic_drift_locs(1500, 128, 160, n_clusters=60, locs_per_frame=40.0, seed=8)
    bounds = np.linspace(0, 1499, 16, dtype=np.uint32)
    dy, dx = imageprocess._rcc_of_locs(locs, info, bounds, 1, 32, lambda i: None)
    hy, hx = imageprocess._rcc_of_locs_windows(locs, info, bounds, 1, 32, lambda i: None)
    np.testing.assert_allclose(dy, hy, atol=1e-5)
    np.testing.assert_allclose(dx, hx, atol=1e-5)
    # a pair subset returns the same per-pair shifts as the full run
    pi, pj = np.triu_indices(15, 1)
    sy, sx = imageprocess._shifts_of_locs(locs, info, bounds, 1, 32)
    sub = slice(3, None, 4)
    sy2, sx2 = imageprocess._shifts_of_locs(locs, info, bounds, 1, 32, pairs=(pi[sub], pj[sub]))
    np.testing.assert_allclose(sy2, sy[sub], atol=1e-5)
    np.testing.assert_allclose(sx2, sx[sub], atol=1e-5)
    # raw records: status 0 almost everywhere, window values equal the window path
    l = _lib.load()
    seg_start, x, y, lpx, lpy = imageprocess._segment_arrays(locs, info, bounds)
    Y_, X_, H, W = imageprocess._crop_geometry(128, 160, 32)
    rec = np.zeros((len(pi), 32)); sums = np.zeros(15)
    win = np.zeros((len(pi), H, W), np.float32)
    pi32, pj32 = pi.astype(np.int32), pj.astype(np.int32)
    _lib.check(l.pb_undrift_peaks_pairs(15, _lib.ptr(seg_start), _lib.ptr(x), _lib.ptr(y), _lib.ptr(lpx),
                                        _lib.ptr(lpy), 128, 160, 1.0, Y_, X_, H, W, len(pi), _lib.ptr(pi32),
                                        _lib.ptr(pj32), _lib.ptr(rec), _lib.ptr(sums), _lib.ptr(win)))
    assert (rec[:, 0] == 0).mean() > 0.9
    for k in (0, 17, 60):
        ym, xm = np.unravel_index(win[k].argmax(), win[k].shape)
        assert (rec[k, 1], rec[k, 2]) == (ym, xm)
        if rec[k, 0] in (0, 1):
            np.testing.assert_array_equal(rec[k, 5:30].reshape(5, 5), win[k][ym - 2:ym + 3, xm - 2:xm + 3])
            xc, yc = imageprocess._gauss_peak_fit(win[k][ym - 2:ym + 3, xm - 2:xm + 3].astype(np.float64))
            assert abs(xc - rec[k, 3]) < 1e-5 and abs(yc - rec[k, 4]) < 1e-5


def test_device_peak_fit_edge_cases():
    """Empty segments (zero sum) and peaks at the crop edge give (0, 0) like the reference."""
    locs, info, _ = testing.synthetic_drift_locs(600, 64, 64, n_clusters=20, locs_per_frame=20.0, seed=2)
    bounds = np.linspace(0, 599, 7, dtype=np.uint32)                      # [0, 99, 199, 299, ...]
    locs = locs[(locs["frame"] < 199) | (locs["frame"] >= 299)].reset_index(drop=True)     # segment 2 empty
    dy, dx = imageprocess._rcc_of_locs(locs, info, bounds, 1, 32, lambda i: None)
    hy, hx = imageprocess._rcc_of_locs_windows(locs, info, bounds, 1, 32, lambda i: None)
    np.testing.assert_allclose(dy, hy, atol=1e-5)
    np.testing.assert_allclose(dx, hx, atol=1e-5)
    sy, sx = imageprocess._shifts_of_locs(locs, info, bounds, 1, 32)
    pi, pj = np.triu_indices(6, 1)
    assert (sy[(pi == 2) | (pj == 2)] == 0).all() and (sx[(pi == 2) | (pj == 2)] == 0).all()


def test_config5_reduced_instance_golden(golden_dir):
    """Reduced BASELINE config 5 (20 segments x 1024^2, 190 pairs; the golden
    per-pair shifts, segment shifts and drift come from the REAL reference run by
    tools/gen_golden.py undrift_c5).  1024^2 images with a 32 x 32 window take the production
    path of the full-size configuration: device render of the segments, batched R2C, the
    FFT-structured pruned inverse transform, device peak fits.  Bar: 1e-3 px."""
    g = np.load(os.path.join(golden_dir, "undrift_c5.npz"))
    locs = pd.DataFrame({k: g[k] for k in ("frame", "x", "y", "lpx", "lpy")})
    H, W, F = (int(v) for v in g["info_hwf"])
    info = [{"Height": H, "Width": W, "Frames": F, "Pixelsize": 130}]
    bounds = g["bounds"]
    n_seg = len(bounds) - 1
    assert n_seg == 20
    sy, sx = imageprocess._shifts_of_locs(locs, info, bounds, 1, 32)
    pi, pj = np.triu_indices(n_seg, 1)
    np.testing.assert_allclose(sy, g["pair_shift_y"][pi, pj], atol=1e-3)
    np.testing.assert_allclose(sx, g["pair_shift_x"][pi, pj], atol=1e-3)
    # the window path (host peak fits) agrees as well
    ry, rx = imageprocess._rcc_of_locs_windows(locs, info, bounds, 1, 32, lambda i: None)
    np.testing.assert_allclose(ry, g["rcc_shift_y"], atol=1e-3)
    np.testing.assert_allclose(rx, g["rcc_shift_x"], atol=1e-3)
    drift, und = postprocess.undrift(locs, info, 100, display=False,
                                     segmentation_callback=lambda i: None, rcc_callback=lambda i: None)
    np.testing.assert_allclose(drift["x"].to_numpy(), g["drift_x"], atol=1e-3)
    np.testing.assert_allclose(drift["y"].to_numpy(), g["drift_y"], atol=1e-3)
