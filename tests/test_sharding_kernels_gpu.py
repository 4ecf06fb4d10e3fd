"""Single-GPU tests of the kernels behind the scalable multi-GPU sharding: the ranks of a world
are simulated one after the other on one device, so the row-band renderer, the band bucketing,
the device-resident fused localize and the device-level undrift pipeline are covered by the
driver's 1-GPU run (the NCCL exchanges themselves: tests/test_multi_gpu.py, gloo tests on CPU)."""
import ctypes as C
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import _lib, distributed as pbd, imageprocess, localize, render as pbrender, testing

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore::DeprecationWarning")]


@pytest.fixture(scope="module")
def torch():
    import torch as t

    return t


@pytest.mark.parametrize("bm", ["gaussian", "gaussian_iso", None])
@pytest.mark.parametrize("world", [1, 3, 8])
def test_row_bands_concatenate_to_the_full_render(torch, bm, world):
    """pb_render_band_count_dev / _scatter_dev / pb_render_band_dev for every band of a `world`:
    bucket ALL localizations by band, render each band from its bucket -> the bands concatenate to
    the single render (n exact, histogram exact, blurred pixels to summation order)."""
    rng = np.random.default_rng(5)
    n = 400_000
    locs = pd.DataFrame({"x": rng.uniform(-1, 49, n).astype(np.float32), "y": rng.uniform(-1, 65, n).astype(np.float32),
                         "lpx": rng.uniform(0.02, 0.25, n).astype(np.float32),
                         "lpy": rng.uniform(0.02, 0.25, n).astype(np.float32)})
    info = [{"Height": 64, "Width": 48, "Frames": 1, "Pixelsize": 130}]
    k1, img1 = pbrender.render(locs, info, oversampling=20, blur_method=bm)
    l = _lib.load()
    pbd._render_decl(l)
    dev = torch.device("cuda", 0)
    t = lambda c: torch.from_numpy(locs[c].to_numpy()).to(dev)      # noqa: E731
    x, y, lpx, lpy = t("x"), t("y"), t("lpx"), t("lpy")
    mode = pbrender._MODES[bm]
    npy, npx = img1.shape
    rows = pbd.band_rows(npy, world)
    args = (20.0, 0.0, 0.0, 64.0, 48.0, 0.0, mode)
    rows_c = (C.c_int * (world + 1))(*rows)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    p = lambda a: a.data_ptr() if a is not None else None           # noqa: E731
    _lib.check(l.pb_render_band_count_dev(n, p(x), p(y), p(lpx), p(lpy), *args, npy, npx, world, rows_c,
                                          counts.data_ptr(), None))
    cnt = counts.cpu().numpy()
    assert cnt.sum() >= k1                                           # halo localizations go to two bands
    offs = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    offsets = torch.from_numpy(offs).to(dev)
    cursor = torch.zeros(world, dtype=torch.int64, device=dev)
    rec = torch.empty((int(cnt.sum()) + 1, 4), dtype=torch.float32, device=dev)
    _lib.check(l.pb_render_band_scatter_dev(n, p(x), p(y), p(lpx), p(lpy), *args, npy, npx, world, rows_c,
                                            offsets.data_ptr(), cursor.data_ptr(), rec.data_ptr(), None))
    np.testing.assert_array_equal(cursor.cpu().numpy(), cnt)
    send = [torch.empty(int(cnt.sum()) + 1, dtype=torch.float32, device=dev) for _ in range(4)]
    _lib.check(l.pb_render_unpack_records_dev(int(cnt.sum()), rec.data_ptr(), *[s.data_ptr() for s in send], None))
    # every in-view localization was written exactly once per band it reaches: the multiset of records of
    # all bands contains each in-view localization at least once
    got = np.sort(send[0][: int(cnt.sum())].cpu().numpy())
    xin = locs["x"].to_numpy()[(locs["x"] > 0) & (locs["y"] > 0) & (locs["x"] < 48) & (locs["y"] < 64)]
    assert len(got) >= len(xin) and np.isin(xin, got).all()
    bands, k_sum = [], 0
    for b in range(world):
        sl = slice(int(offs[b]), int(offs[b] + cnt[b]))
        nb = int(cnt[b])
        h = rows[b + 1] - rows[b]
        img = torch.empty((h, npx), dtype=torch.float32, device=dev)
        c = torch.zeros(1, dtype=torch.int64, device=dev)
        wsb = int(l.pb_render_workspace_bytes(nb, max(h, 1), npx))
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        bx, by, blx, bly = (s[sl].contiguous() for s in send)
        _lib.check(l.pb_render_band_dev(nb, bx.data_ptr() if nb else None, by.data_ptr() if nb else None,
                                        blx.data_ptr() if nb else None, bly.data_ptr() if nb else None, *args,
                                        img.data_ptr() if h else None, npy, npx, rows[b], h, c.data_ptr(),
                                        ws.data_ptr(), wsb, None))
        torch.cuda.synchronize()
        bands.append(img.cpu().numpy())
        k_sum += int(c.item())
    full = np.concatenate(bands, 0)
    assert k_sum == k1
    assert full.shape == img1.shape
    if bm is None:
        np.testing.assert_array_equal(full, img1)
    else:
        big = img1 > 1e-3 * img1.max()
        np.testing.assert_allclose(full[big], img1[big], rtol=1e-4)
        np.testing.assert_allclose(full, img1, rtol=1e-4, atol=1e-6 * float(img1.max()))


def test_render_bands_device_single_rank_equals_render(torch):
    rng = np.random.default_rng(6)
    n = 300_000
    locs = pd.DataFrame({"x": rng.uniform(0, 64, n).astype(np.float32), "y": rng.uniform(0, 48, n).astype(np.float32),
                         "lpx": rng.uniform(0.02, 0.08, n).astype(np.float32),
                         "lpy": rng.uniform(0.02, 0.08, n).astype(np.float32)})
    info = [{"Height": 48, "Width": 64, "Frames": 1, "Pixelsize": 130}]
    k1, img1 = pbrender.render(locs, info, oversampling=20, blur_method="gaussian")
    k, band, rr = pbd.render_bands(None, torch, locs, info, device="cuda:0", oversampling=20, blur_method="gaussian")
    assert k == k1 and rr == (0, 960) and band.shape == img1.shape
    big = img1 > 1e-3 * img1.max()
    np.testing.assert_allclose(band[big], img1[big], rtol=1e-4)


def test_localize_device_equals_host_path(torch):
    movie = testing.synthetic_movie(24, 96, 80, emitters_per_frame=12, seed=4)
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    par = {"Min. Net Gradient": 5000, "Box Size": 7}
    dm = torch.from_numpy(movie.view(np.int16)).to("cuda:0")
    for method, names in (("gausslq", localize.LOCS_COLUMNS_LQ), ("gaussmle", localize.LOCS_COLUMNS_MLE)):
        ref = localize.localize(movie, cam, par, fitting_method=method, return_info=False)
        cols = pbd.localize_device(torch, dm, 0, cam, par, fitting_method=method).cpu().numpy()
        got = localize._columns_to_locs(np.ascontiguousarray(cols), names)
        assert len(got) == len(ref) and len(ref) > 100
        for c in ref.columns:
            assert got[c].to_numpy().tobytes() == ref[c].to_numpy().tobytes(), c
        # a frame block with an offset: frame numbers shift, everything else is the block's own
        blk = pbd.localize_device(torch, dm[8:16], 8, cam, par, fitting_method=method).cpu().numpy()
        sub = ref[(ref["frame"] >= 8) & (ref["frame"] < 16)]
        gb = localize._columns_to_locs(np.ascontiguousarray(blk), names)
        assert len(gb) == len(sub)
        for c in ref.columns:
            assert gb[c].to_numpy().tobytes() == sub[c].to_numpy().tobytes(), c


def test_undrift_shifts_device_single_rank_golden(torch, golden_dir):
    """The device-level undrift pipeline (render -> R2C -> tile-sorted pairs -> peak fits) on the
    reduced config-5 golden: per-pair shifts within 1e-3 px of the REAL reference and equal to the
    monolithic pb_undrift_peaks_pairs path."""
    g = np.load(os.path.join(golden_dir, "undrift_c5.npz"))
    locs = pd.DataFrame({k: g[k] for k in ("frame", "x", "y", "lpx", "lpy")})
    H, W, F = (int(v) for v in g["info_hwf"])
    info = [{"Height": H, "Width": W, "Frames": F, "Pixelsize": 130}]
    bounds = g["bounds"]
    n_seg = len(bounds) - 1
    seg_start, x, y, lpx, lpy = imageprocess._segment_arrays(locs, info, bounds)
    t = lambda a: torch.from_numpy(a).to("cuda:0")      # noqa: E731
    timings = {}
    sy, sx = pbd.undrift_shifts_device(None, torch, seg_start, t(x), t(y), t(lpx), t(lpy), n_seg, H, W,
                                       timings=timings)
    pi, pj = np.triu_indices(n_seg, 1)
    np.testing.assert_allclose(sy, g["pair_shift_y"][pi, pj], atol=1e-3)
    np.testing.assert_allclose(sx, g["pair_shift_x"][pi, pj], atol=1e-3)
    my, mx = imageprocess._shifts_of_locs(locs, info, bounds, 1, 32)
    np.testing.assert_allclose(sy, my, atol=2e-6)
    np.testing.assert_allclose(sx, mx, atol=2e-6)
    assert {"render_ms", "r2c_ms", "pairs_ms", "peakfit_ms"} <= set(timings)
    drift, und = pbd.undrift_segments_sharded(None, torch, locs, info, 100, device="cuda:0")
    np.testing.assert_allclose(drift["x"].to_numpy(), g["drift_x"], atol=1e-3)
    np.testing.assert_allclose(drift["y"].to_numpy(), g["drift_y"], atol=1e-3)
