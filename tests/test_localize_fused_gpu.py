"""GPU parity of the fused movie -> localization-table pipeline (csrc/localize.cu, SURVEY.md
section 8f ranks 1-2): the device-side column arithmetic must be BIT-IDENTICAL to the real
reference's locs_from_fits DataFrames (golden), and pb_localize must give exactly the table of
identify -> get_spots -> fit -> locs_from_fits."""
import ctypes as C
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import _lib, gausslq, gaussmle, localize, testing

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore::DeprecationWarning")]

CAM = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "testdata.npz"))


@pytest.fixture(scope="module")
def ids(g):
    return pd.DataFrame({"frame": g["ids_frame"], "x": g["ids_x"], "y": g["ids_y"],
                         "net_gradient": g["ids_ng"]})


def _sorted_like_reference(cols):
    locs = pd.DataFrame(cols)
    locs.sort_values(by="frame", kind="quicksort", inplace=True)
    return locs


def _assert_bits(locs, g, prefix):
    names = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
    assert list(locs.columns) == names
    for c in names:
        ref = g[prefix + c]
        got = locs[c].to_numpy()
        assert got.dtype == ref.dtype, c
        assert got.tobytes() == ref.tobytes(), c


def test_device_columns_mle_bit_identical_to_reference(g, ids):
    cols = localize.locs_columns_from_fits(ids, g["mle_thetas"], 7, 1, CRLBs=g["mle_crlbs"],
                                           log_likelihoods=g["mle_logliks"],
                                           iterations=g["mle_iterations"])
    _assert_bits(_sorted_like_reference(cols), g, "locs_")


@pytest.mark.parametrize("em", [False, True])
def test_device_columns_lq_bit_identical_to_reference(g, ids, em):
    cols = localize.locs_columns_from_fits(ids, g["lq_thetas"], 7, 2, em=em)
    _assert_bits(_sorted_like_reference(cols), g, f"lqlocs_em{int(em)}_")


def test_device_columns_gpufit_layout_bit_identical_to_reference(g, ids):
    lq = g["lq_thetas"]
    gp = np.stack([lq[:, 2], lq[:, 0] + 3, lq[:, 1] + 3, lq[:, 4], lq[:, 5], lq[:, 3]], 1)
    cols = localize.locs_columns_from_fits(ids, gp, 7, 3)
    locs = _sorted_like_reference(cols)
    for c in (k[len("gplocs_"):] for k in g.files if k.startswith("gplocs_")):
        assert locs[c].to_numpy().tobytes() == g[f"gplocs_{c}"].tobytes(), c


def test_device_columns_special_values():
    """NaN / inf / zero sigma propagate like numpy (np.maximum keeps NaN, sqrt(inf) = inf)."""
    idf = pd.DataFrame({"frame": [0, 1, 2], "x": [3, 4, 5], "y": [6, 7, 8],
                        "net_gradient": np.float32([1, 2, 3])})
    th = np.float32([[3.2, 2.9, 1000, 10, 1.1, np.nan], [3, 3, 500, 5, 0.0, 0.0],
                     [2.5, 3.5, 800, 0.01, 1.3, 0.9]])
    cr = np.float32([[np.inf] * 6, [0, 1e-4, 4, 9, 16, 25], [1e-4, 4e-4, 1, -1, 2, 3]])
    ll = np.float32([-30, -40, np.nan])
    it = np.int32([100, 3, 0])
    with np.errstate(all="ignore"):
        ref = gaussmle.locs_from_fits(idf, th, cr, ll, it, 7)
        cols = localize.locs_columns_from_fits(idf, th, 7, 0, CRLBs=cr, log_likelihoods=ll, iterations=it)
        for c in ref.columns:
            np.testing.assert_array_equal(cols[c], ref[c].to_numpy(), err_msg=c)
        ref = gausslq.locs_from_fits(idf, th, 7, True)
        cols = localize.locs_columns_from_fits(idf, th, 7, 2, em=True)
        for c in ref.columns:
            np.testing.assert_array_equal(cols[c], ref[c].to_numpy(), err_msg=c)


def _two_call(movie, method, mle_method="sigmaxy", roi=None, frame_bounds=None, cam=CAM, box=7, mng=5000):
    idf = localize.identify(movie, mng, box, roi=roi, frame_bounds=frame_bounds, return_info=False)
    locs, _ = localize.fit2D(movie, [], dict(cam), idf, box, fitting_method=method, mle_method=mle_method)
    return idf, locs


@pytest.mark.parametrize("method,mle_method", [("gaussmle", "sigmaxy"), ("gaussmle", "sigma"),
                                               ("gausslq", "sigmaxy"), ("gausslq-gpu", "sigmaxy")])
@pytest.mark.parametrize("chunk_frames", [None, 3])
def test_fused_equals_two_call_path(monkeypatch, method, mle_method, chunk_frames):
    """Several device chunks (PB_LOCALIZE_CHUNK_FRAMES) and a single one give the table of the
    identify -> fit2D path bit for bit, EM camera included."""
    if chunk_frames:
        monkeypatch.setenv("PB_LOCALIZE_CHUNK_FRAMES", str(chunk_frames))
    cam = {"Baseline": 90, "Sensitivity": 0.45, "Gain": 2, "Pixelsize": 130}
    movie = testing.synthetic_movie(14, 72, 88, emitters_per_frame=7, seed=33)
    locs = localize.localize(movie, dict(cam), {"Min. Net Gradient": 4000, "Box Size": 7},
                             fitting_method=method, mle_method=mle_method, return_info=False)
    idf, locs2 = _two_call(movie, method, mle_method, cam=cam, mng=4000)
    assert len(locs) == len(idf) > 40
    pd.testing.assert_frame_equal(locs.reset_index(drop=True), locs2.reset_index(drop=True),
                                  check_exact=True)


@pytest.mark.parametrize("box", [5, 9, 13])
def test_fused_other_boxes_float_movie_roi_bounds(box):
    movie = testing.synthetic_movie(10, 90, 90, emitters_per_frame=6, seed=5).astype(np.float32)
    roi = ((10, 5), (80, 85))
    fb = (2, 7)
    seen = []
    locs, info = localize.localize(movie, dict(CAM), {"Min. Net Gradient": 3000, "Box Size": box}, roi=roi,
                                   frame_bounds=fb, fitting_method="gaussmle", return_info=True,
                                   identification_progress_callback=seen.append)
    idf, locs2 = _two_call(movie, "gaussmle", roi=roi, frame_bounds=fb, box=box, mng=3000)
    assert seen and seen[-1] == len(movie)
    assert len(locs) == len(idf) > 10
    assert locs["frame"].min() >= 2 and locs["frame"].max() <= 7
    pd.testing.assert_frame_equal(locs.reset_index(drop=True), locs2.reset_index(drop=True),
                                  check_exact=True)
    assert info[0]["ROI"] == roi and info[0]["Frame Bounds"] == fb


def test_fused_bundled_movie_known_answers(g):
    """The reference's tests/data/testdata.raw: 30 identifications; fused MLE / LQ tables equal the
    real reference's DataFrames within the fit parity bars (identification columns bit-exact)."""
    movie = g["movie"]
    cam = {"Baseline": 0, "Sensitivity": 1, "Gain": 1, "Pixelsize": 130}
    locs = localize.localize(movie, dict(cam), {"Min. Net Gradient": 5000, "Box Size": 7},
                             fitting_method="gaussmle", return_info=False)
    assert len(locs) == 30
    np.testing.assert_array_equal(locs["frame"].to_numpy(), g["locs_frame"])
    assert locs["net_gradient"].to_numpy().tobytes() == g["locs_net_gradient"].tobytes()
    for c, tol in (("x", 1e-4), ("y", 1e-4), ("sx", 1e-4), ("sy", 1e-4), ("lpx", 1e-5), ("lpy", 1e-5)):
        np.testing.assert_allclose(locs[c].to_numpy(), g[f"locs_{c}"], atol=tol, rtol=0, err_msg=c)
    np.testing.assert_allclose(locs["photons"].to_numpy(), g["locs_photons"], rtol=1e-4)
    np.testing.assert_allclose(locs["log_likelihood"].to_numpy(), g["locs_log_likelihood"], rtol=1e-4)
    lq = localize.localize(movie, dict(cam), {"Min. Net Gradient": 5000, "Box Size": 7},
                           fitting_method="gausslq", return_info=False)
    assert len(lq) == 30
    for c in ("x", "y", "sx", "sy", "lpx", "lpy"):
        np.testing.assert_allclose(lq[c].to_numpy(), g[f"lqlocs_em0_{c}"], atol=2e-4, rtol=0, err_msg=c)


def test_pb_localize_capacity_protocol_and_errors():
    lib = _lib.load()
    localize._declare(lib)
    movie = testing.synthetic_movie(6, 64, 64, emitters_per_frame=5, seed=9)
    found = C.c_size_t(0)
    cols = np.zeros((17, 4), np.float32)
    args = lambda cap, fit=1, box=7: (_lib.ptr(movie), 0, 6, 64, 64, 0, box, 5000.0, None, 100.0, 1.0, 1.0,  # noqa: E731
                                      fit, 1e-3, 100, 0, _lib.ptr(cols), cap, C.byref(found))
    assert lib.pb_localize(*args(4)) == 4 and found.value > 4          # PB_ERR_CAPACITY + required size
    need = found.value
    cols = np.zeros((17, need), np.float32)
    assert lib.pb_localize(*args(need)) == 0 and found.value == need
    fr = cols[0].view(np.uint32)
    assert (np.diff(fr.astype(np.int64)) >= 0).all()                    # rows ordered by frame
    assert lib.pb_localize(*args(need, fit=7)) == 1
    assert lib.pb_localize(*args(need, box=4)) == 1
    assert lib.pb_locs_columns(1) == 17 and lib.pb_locs_columns(2) == 11 and lib.pb_locs_columns(9) == -1
    # empty movie
    assert lib.pb_localize(_lib.ptr(movie), 0, 0, 64, 64, 0, 7, 5000.0, None, 100.0, 1.0, 1.0, 1, 1e-3, 100,
                           0, _lib.ptr(cols), need, C.byref(found)) == 0 and found.value == 0


def test_fused_pinned_movie_and_no_spots():
    """A pinned movie (pb_host_alloc) is copied without staging; a movie without spots gives an
    empty table with the reference's columns."""
    movie = testing.synthetic_movie(5, 64, 64, emitters_per_frame=4, seed=2)
    pin = _lib.PinnedArray(movie.shape, movie.dtype)
    pin.array[...] = movie
    a = localize.localize(pin.array, dict(CAM), {"Min. Net Gradient": 5000, "Box Size": 7},
                          fitting_method="gaussmle", return_info=False)
    b = localize.localize(movie, dict(CAM), {"Min. Net Gradient": 5000, "Box Size": 7},
                          fitting_method="gaussmle", return_info=False)
    pd.testing.assert_frame_equal(a.reset_index(drop=True), b.reset_index(drop=True), check_exact=True)
    pin.free()
    flat = np.full((3, 40, 40), 100, np.uint16)
    e = localize.localize(flat, dict(CAM), {"Min. Net Gradient": 5000, "Box Size": 7},
                          fitting_method="gausslq", return_info=False)
    assert len(e) == 0 and list(e.columns) == list(localize.LOCS_COLUMNS_LQ)
