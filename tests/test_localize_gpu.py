"""GPU parity: identify / get_spots through the C ABI vs golden vectors from the real
reference and vs the CPU oracle.  Integer results and float32 net gradients must be
BIT-EXACT (coordinates, counts, order of the serial reference)."""
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import localize, testing

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "identify.npz"))


@pytest.mark.parametrize("box", [3, 5, 7, 9, 11, 13])
def test_tie_rule_golden(gold, box):
    # min_ng = -inf keeps every local maximum -> pins the argmax tie rule
    y, x, ng = localize.identify_in_image(gold[f"ties_b{box}_frame"], -1e30, box)
    np.testing.assert_array_equal(y, gold[f"ties_b{box}_y"])
    np.testing.assert_array_equal(x, gold[f"ties_b{box}_x"])


@pytest.mark.parametrize("box", [5, 7, 9])
def test_wraparound_golden(gold, box):
    y, x, ng = localize.identify_in_image(gold[f"wrap_b{box}_frame"], -1e30, box)
    np.testing.assert_array_equal(y, gold[f"wrap_b{box}_y"])
    np.testing.assert_array_equal(x, gold[f"wrap_b{box}_x"])
    assert (_bits(ng) == _bits(gold[f"wrap_b{box}_ng"])).all()


@pytest.mark.parametrize("box,mng", [(7, 5000), (9, 8000), (5, 3000)])
def test_movie_golden(gold, box, mng):
    ids = localize.identify(gold["movie"], mng, box, return_info=False)
    tag = f"mov_b{box}"
    assert ids["frame"].dtype == np.int64 and ids["net_gradient"].dtype == np.float32
    np.testing.assert_array_equal(ids["frame"].to_numpy(), gold[f"{tag}_frame"])
    np.testing.assert_array_equal(ids["x"].to_numpy(), gold[f"{tag}_x"])
    np.testing.assert_array_equal(ids["y"].to_numpy(), gold[f"{tag}_y"])
    assert (_bits(ids["net_gradient"].to_numpy()) == _bits(gold[f"{tag}_ng"])).all()
    cam = {"Baseline": 100, "Sensitivity": 0.45, "Gain": 2}
    spots = localize.get_spots(gold["movie"], ids, box, cam)
    assert spots.dtype == np.float32 and spots.shape == (len(ids), box, box)
    assert (_bits(spots) == _bits(gold[f"{tag}_spots"])).all()


def test_roi_and_frame_bounds_golden(gold):
    roi = tuple(map(tuple, gold["roi"].tolist()))
    fb = tuple(gold["roi_frame_bounds"].tolist())
    ids, info = localize.identify(gold["movie"], 5000, 7, roi=roi, frame_bounds=fb,
                                  return_info=True)
    np.testing.assert_array_equal(ids["frame"].to_numpy(), gold["roi_frame"])
    np.testing.assert_array_equal(ids["x"].to_numpy(), gold["roi_x"])
    np.testing.assert_array_equal(ids["y"].to_numpy(), gold["roi_y"])
    assert (_bits(ids["net_gradient"].to_numpy()) == _bits(gold["roi_ng"])).all()
    assert info["Box Size"] == 7 and info["ROI"] == roi


def test_bundled_testdata(golden_dir):
    g = np.load(os.path.join(golden_dir, "testdata.npz"))
    ids = localize.identify(g["movie"], 5000, 7, return_info=False)
    assert len(ids) == 30
    np.testing.assert_array_equal(ids["frame"].to_numpy(), g["ids_frame"])
    np.testing.assert_array_equal(ids["x"].to_numpy(), g["ids_x"])
    np.testing.assert_array_equal(ids["y"].to_numpy(), g["ids_y"])
    assert (_bits(ids["net_gradient"].to_numpy()) == _bits(g["ids_ng"])).all()
    spots = localize.get_spots(g["movie"], ids, 7, {"Baseline": 0, "Sensitivity": 1, "Gain": 1})
    np.testing.assert_array_equal(spots, g["spots"])


@pytest.mark.parametrize("shape,box", [((6, 200, 333), 7), ((3, 512, 512), 7), ((4, 97, 1031), 9),
                                       ((2, 40, 20), 5), ((5, 130, 260), 11), ((2, 16, 16), 15)])
def test_identify_vs_oracle_random(oracle, shape, box):
    """Odd widths (scalar load path), multi-tile frames, tiny frames, heavy ties."""
    F, Y, X = shape
    movie = testing.synthetic_movie(F, Y, X, emitters_per_frame=max(2, Y * X // 4000), seed=F + box,
                                    margin=min(8, Y // 4))
    movie[0] = (movie[0] // 16) * 16            # plateaus -> ties
    movie[-1, -1, :] += 3000                    # bright last row (wrap-around source)
    for mng in (-1e30, 1000.0):
        ofr, ox, oy, ong = oracle.identify_movie(movie, mng, box)
        ids = localize.identify(movie, mng, box, return_info=False)
        np.testing.assert_array_equal(ids["frame"].to_numpy(), ofr)
        np.testing.assert_array_equal(ids["y"].to_numpy(), oy)
        np.testing.assert_array_equal(ids["x"].to_numpy(), ox)
        assert (_bits(ids["net_gradient"].to_numpy()) == _bits(ong)).all()
    spots = localize.get_spots(movie, ids, box, {"Baseline": 97.5, "Sensitivity": 0.31, "Gain": 3})
    ospots = oracle.get_spots(movie, ofr, ox, oy, box, {"Baseline": 97.5, "Sensitivity": 0.31,
                                                       "Gain": 3})
    assert (_bits(spots) == _bits(ospots)).all()


def test_identify_float_image_and_frame_api(oracle):
    movie = testing.synthetic_movie(2, 64, 64, emitters_per_frame=6, seed=11)
    img = movie[0].astype(np.float32) * 0.37
    y, x, ng = localize.identify_in_image(img, 100.0, 7)
    oy, ox, ong = oracle.identify_in_image(img, 100.0, 7)
    np.testing.assert_array_equal(y, oy)
    np.testing.assert_array_equal(x, ox)
    assert (_bits(ng) == _bits(ong)).all()
    df = localize.identify_by_frame_number(movie, 5000, 7, 1)
    assert list(df.columns) == ["frame", "x", "y", "net_gradient"] and (df["frame"] == 1).all()
    df = localize.identify_by_frame_number(movie, 5000, 7, 1, frame_bounds=(0, 0))
    assert len(df) == 0
    # flat frame -> nothing
    y, x, ng = localize.identify_in_frame(np.full((32, 32), 7, np.uint16), 0, 7)
    assert len(y) == 0


def test_progress_abort_and_empty():
    movie = testing.synthetic_movie(5, 48, 48, emitters_per_frame=3, seed=2)
    seen = []
    ids = localize.identify(movie, 5000, 7, threaded=False, progress_callback=seen.append,
                            return_info=False)
    assert seen == list(range(5))
    assert localize.identify(movie, 5000, 7, abort_callback=lambda: True, return_info=False) is None
    empty = localize.get_spots(movie, ids.iloc[:0], 7, {"Baseline": 0, "Sensitivity": 1, "Gain": 1})
    assert empty.shape == (0, 7, 7)
    with pytest.warns(DeprecationWarning):
        localize.identify(movie, 5000, 7)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("method", ["gausslq", "gaussmle", "gausslq-gpu"])
def test_localize_equals_identify_plus_fit2d(method):
    """Reference test_localize.py:849-1069: localize == identify + fit2D; metadata keys;
    the fused one-upload path gives the same table as the two-call path."""
    movie = testing.synthetic_movie(16, 96, 80, emitters_per_frame=8, seed=21)
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    params = {"Min. Net Gradient": 5000, "Box Size": 7}
    locs, info = localize.localize(movie, dict(cam), params, fitting_method=method,
                                   return_info=True, movie_info=[{"Frames": 16}])
    ids = localize.identify(movie, 5000, 7, return_info=False)
    locs2, fit_info = localize.fit2D(movie, [{"Frames": 16}], dict(cam), ids, 7,
                                     fitting_method=method)
    assert len(locs) == len(ids) == len(locs2) > 50
    pd.testing.assert_frame_equal(locs.reset_index(drop=True), locs2.reset_index(drop=True))
    assert info[0] == {"Frames": 16} and info[1]["Box Size"] == 7
    assert info[2]["Fit method"] == method and info[2]["Pixelsize"] == 130
    if method == "gaussmle":
        assert info[2]["Convergence criterion"] == 0.001 and info[2]["Max iterations"] == 100
        assert list(locs.columns) == ["frame", "x", "y", "photons", "sx", "sy", "bg", "lpx", "lpy",
                                      "ellipticity", "net_gradient", "log_likelihood", "iterations",
                                      "photons_unc", "bg_unc", "sx_unc", "sy_unc"]
        assert locs["frame"].dtype == np.uint32 and locs["iterations"].dtype == np.uint32
    # positions land near the identification pixel
    assert (np.abs(locs["x"].to_numpy() - ids["x"].to_numpy()) < 2).all()


def test_fit2d_argument_contract():
    movie = testing.synthetic_movie(2, 40, 40, emitters_per_frame=2, seed=3)
    ids = localize.identify(movie, 5000, 7, return_info=False)
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1}
    with pytest.warns(UserWarning, match="Pixelsize"):
        locs, info = localize.fit2D(movie, [], cam, ids, 7)
    assert cam["Pixelsize"] == 130 and len(locs) == len(ids)
    with pytest.raises(AssertionError):
        localize.fit2D(movie, [], cam, ids, 7, fitting_method="nope")
    with pytest.raises(AssertionError):
        localize.fit2D(movie, [], cam, ids, 7, eps=-1)
    assert localize.fit2D(movie, [], cam, ids, 7, abort_callback=lambda: True)[0] is None
    with pytest.raises(NotImplementedError):
        localize.fit2D(movie, [], cam, ids, 7, fitting_method="avg")


def test_identify_async_matches_identify():
    import time

    movie = testing.synthetic_movie(9, 64, 64, emitters_per_frame=5, seed=8)
    ids = localize.identify(movie, 5000, 7, return_info=False)
    current, futures = localize.identify_async(movie, 5000, 7)
    t0 = time.time()
    while current[0] < len(movie) and time.time() - t0 < 60:
        time.sleep(0.01)
    assert current[0] == len(movie)
    ids2 = localize.identifications_from_futures(futures)
    a = sorted(zip(ids["frame"], ids["y"], ids["x"]))
    b = sorted(zip(ids2["frame"], ids2["y"], ids2["x"]))
    assert a == b and len(a) > 10


def test_config3_full_size_vs_oracle(oracle):
    """BASELINE config 3 at its full size (2000 x 512 x 512 uint16, 60 emitters / frame, box 7,
    min net gradient 5000): identifications equal to the oracle's as sorted
    (frame, y, x) tuples with bit-equal float32 net gradients, then the LQ fit of the cut-out ROIs
    within the LQ tolerance.  The movie is generated in 20 chunks of 100 frames (seeds 1..20) on a
    thread pool, the oracle runs chunk-wise on the same pool."""
    from concurrent.futures import ThreadPoolExecutor

    nchunk, per = 20, 100
    with ThreadPoolExecutor(min(nchunk, os.cpu_count() or 1)) as ex:
        chunks = list(ex.map(lambda k: testing.synthetic_movie(per, 512, 512, emitters_per_frame=60,
                                                               seed=1 + k), range(nchunk)))
        movie = np.concatenate(chunks)
        del chunks
        assert movie.shape == (2000, 512, 512) and movie.dtype == np.uint16
        parts = list(ex.map(lambda k: oracle.identify_movie(movie[k * per:(k + 1) * per], 5000, 7),
                            range(nchunk)))
    ofr = np.concatenate([p[0] + k * per for k, p in enumerate(parts)])
    ox = np.concatenate([p[1] for p in parts]); oy = np.concatenate([p[2] for p in parts])
    ong = np.concatenate([p[3] for p in parts])
    ids = localize.identify(movie, 5000, 7, return_info=False)
    assert len(ids) == len(ofr) and len(ids) > 100_000
    a = np.lexsort((ids["x"].to_numpy(), ids["y"].to_numpy(), ids["frame"].to_numpy()))
    b = np.lexsort((ox, oy, ofr))
    for c, o in (("frame", ofr), ("y", oy), ("x", ox)):
        np.testing.assert_array_equal(ids[c].to_numpy()[a], o[b])
    assert (_bits(ids["net_gradient"].to_numpy()[a]) == _bits(ong[b])).all()
    # ROIs + LQ fit of a 20 k sub-sample against the oracle (bit-exact ROIs, LQ tolerance)
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    sub = ids.iloc[:: max(1, len(ids) // 20000)]
    spots = localize.get_spots(movie, sub, 7, cam)
    ospots = oracle.get_spots(movie, sub["frame"].to_numpy(), sub["x"].to_numpy(), sub["y"].to_numpy(), 7, cam)
    assert (_bits(spots) == _bits(ospots)).all()
    from picasso_b200 import gausslq

    th = gausslq.fit_spots(spots)
    oth = oracle.fit_spots_lq(spots, nthreads=os.cpu_count() or 1)
    d = th.astype(np.float64) - oth
    rms = np.sqrt((d ** 2).mean(0))
    assert rms[[0, 1, 4, 5]].max() <= 1e-4, rms
    # the fused movie -> table path finds the same localizations
    locs = localize.localize(movie, cam, {"Min. Net Gradient": 5000, "Box Size": 7},
                             fitting_method="gausslq", return_info=False)
    assert len(locs) == len(ids)
    np.testing.assert_array_equal(np.sort(locs["frame"].to_numpy()), np.sort(ofr).astype(np.uint32))
