"""Generate the golden vectors under tests/golden/ by running the REAL reference
(jungmannlab/picasso mounted at /root/reference, imported with its GUI/IO
dependencies mocked -- tools/ref_import.py).

Run in the build container only:

    python tools/gen_golden.py [mle] [identify] [lq] [render] [undrift] [undrift_c5] [testdata]

Each fixture stores the inputs and the reference's outputs, so the tests need
neither the reference nor this script at run time.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_mle(ref):
    from picasso_b200 import testing

    gm = ref["gaussmle"]
    # config 1: 10k Poisson spots, box 7 -- counts fit in uint16
    spots = testing.synthetic_spots(10_000, 7, seed=0)
    assert spots.max() < 65535 and (spots == np.round(spots)).all()
    out = {"spots_u16": spots.astype(np.uint16)}
    for method in ("sigmaxy", "sigma"):
        th, cr, ll, it = gm.gaussmle(spots, 0.001, 100, method)
        out[f"{method}_thetas"] = th
        out[f"{method}_crlbs"] = cr
        out[f"{method}_logliks"] = ll
        out[f"{method}_iterations"] = it
    save("mle_config1.npz", **out)

    # other box sizes, 300 spots each, both methods
    out = {}
    for box in (5, 9, 11, 13, 15):
        spots = testing.synthetic_spots(300, box, seed=100 + box)
        out[f"b{box}_spots_u16"] = spots.astype(np.uint16)
        for method in ("sigmaxy", "sigma"):
            th, cr, ll, it = gm.gaussmle(spots, 0.001, 100, method)
            out[f"b{box}_{method}_thetas"] = th
            out[f"b{box}_{method}_crlbs"] = cr
            out[f"b{box}_{method}_logliks"] = ll
            out[f"b{box}_{method}_iterations"] = it
    save("mle_boxes.npz", **out)

    # non-integer (noiseless, point-sampled) float spots + tighter eps / capped
    # iterations: the regimes the reference's own tests use
    # (tests/conftest.py:121-188, tests/test_gaussmle.py:129-140)
    rng = np.random.default_rng(42)
    n, box = 64, 7
    half = box // 2
    grid = np.arange(-half, half + 1, dtype=np.float64)
    x0 = rng.uniform(-0.5, 0.5, n); y0 = rng.uniform(-0.5, 0.5, n)
    sx = rng.uniform(0.8, 1.4, n); sy = rng.uniform(0.8, 1.4, n)
    ph = rng.uniform(2000, 8000, n); bg = rng.uniform(5, 50, n)
    gx = np.exp(-0.5 * ((grid[None] - x0[:, None]) / sx[:, None]) ** 2) / (sx[:, None] * np.sqrt(2 * np.pi))
    gy = np.exp(-0.5 * ((grid[None] - y0[:, None]) / sy[:, None]) ** 2) / (sy[:, None] * np.sqrt(2 * np.pi))
    clean = (ph[:, None, None] * gy[:, :, None] * gx[:, None, :] + bg[:, None, None]).astype(np.float32)
    out = {"spots": clean}
    for tag, eps, max_it in (("e3", 1e-3, 100), ("e6", 1e-6, 100), ("it3", 1e-3, 3), ("it0", 1e-3, 0)):
        for method in ("sigmaxy", "sigma"):
            th, cr, ll, it = gm.gaussmle(clean, eps, max_it, method)
            out[f"{tag}_{method}_thetas"] = th
            out[f"{tag}_{method}_crlbs"] = cr
            out[f"{tag}_{method}_logliks"] = ll
            out[f"{tag}_{method}_iterations"] = it
    save("mle_float_spots.npz", **out)


def gen_identify(ref):
    """identify / get_spots golden vectors (reference localize.py:97-337, 917-1145)."""
    from picasso_b200 import testing

    loc = ref["localize"]
    out = {}
    rng = np.random.default_rng(7)
    # 1) tie-heavy small-integer frames: pins the first-in-row-major argmax rule
    cases = []
    for k, box in enumerate((3, 5, 7, 9, 11, 13)):
        Y, X = 40 + 3 * k, 52 - 2 * k
        fr = rng.integers(0, 2 * box * box, (Y, X)).astype(np.float32)
        y, x = loc._local_maxima(fr, box)
        out[f"ties_b{box}_frame"] = fr
        out[f"ties_b{box}_y"] = y
        out[f"ties_b{box}_x"] = x
        cases.append(len(y))
    # 2) net-gradient wrap-around: bright last row / column, candidates at i == box_half
    for box in (5, 7, 9):
        Y, X = 30, 34
        fr = rng.integers(0, 50, (Y, X)).astype(np.float32)
        fr[-1, :] += 3000
        fr[:, -1] += 2000
        h = box // 2
        fr[h, h] = 5000; fr[h, 17] = 4000; fr[15, h] = 4500; fr[12, 20] = 6000
        y, x, ng = loc.identify_in_image(fr, -1e30, box)
        out[f"wrap_b{box}_frame"] = fr
        out[f"wrap_b{box}_y"] = y; out[f"wrap_b{box}_x"] = x; out[f"wrap_b{box}_ng"] = ng
    # 3) realistic movie: serial identify, roi, frame bounds, get_spots
    movie = testing.synthetic_movie(12, 64, 72, emitters_per_frame=10, seed=5)
    out["movie"] = movie
    for box, mng in ((7, 5000), (9, 8000), (5, 3000)):
        ids = loc._identify_serial(movie, mng, box, None, None, None)
        tag = f"mov_b{box}"
        out[f"{tag}_frame"] = ids["frame"].to_numpy(); out[f"{tag}_x"] = ids["x"].to_numpy()
        out[f"{tag}_y"] = ids["y"].to_numpy(); out[f"{tag}_ng"] = ids["net_gradient"].to_numpy()
        cam = {"Baseline": 100, "Sensitivity": 0.45, "Gain": 2, "Qe": 0.9}
        out[f"{tag}_spots"] = loc.get_spots(movie, ids, box, cam)
    roi = ((10, 12), (50, 61))
    ids = loc._identify_serial(movie, 5000, 7, roi, (3, 8), None)
    out["roi_frame"] = ids["frame"].to_numpy(); out["roi_x"] = ids["x"].to_numpy()
    out["roi_y"] = ids["y"].to_numpy(); out["roi_ng"] = ids["net_gradient"].to_numpy()
    out["roi"] = np.array(roi); out["roi_frame_bounds"] = np.array([3, 8])
    save("identify.npz", **out)
    print("tie cases maxima counts:", cases)


def gen_testdata(ref):
    """Known answers on the reference's bundled test movie (tests/data/testdata.raw,
    100x32x32 uint16; SURVEY.md 8c 'verified oracle outputs')."""
    loc, gm, glq, rnd = ref["localize"], ref["gaussmle"], ref["gausslq"], ref["render"]
    import pandas as pd

    movie = np.fromfile(os.path.join(ref_import.REFERENCE_ROOT, "tests", "data", "testdata.raw"),
                        dtype="<u2").reshape(100, 32, 32)
    ids = loc._identify_serial(movie, 5000, 7, None, None, None)
    cam = {"Baseline": 0, "Sensitivity": 1, "Gain": 1}
    spots = loc.get_spots(movie, ids, 7, cam)
    th, cr, ll, it = gm.gaussmle(spots, 0.001, 100, "sigmaxy")
    ths, crs, lls, its = gm.gaussmle(spots, 0.001, 100, "sigma")
    lq = glq.fit_spots(spots)
    locs = gm.locs_from_fits(ids, th, cr, ll, it, 7)
    info = [{"Height": 32, "Width": 32, "Frames": 100, "Pixelsize": 130}]
    out = dict(movie=movie, ids_frame=ids["frame"].to_numpy(), ids_x=ids["x"].to_numpy(),
               ids_y=ids["y"].to_numpy(), ids_ng=ids["net_gradient"].to_numpy(), spots=spots,
               mle_thetas=th, mle_crlbs=cr, mle_logliks=ll, mle_iterations=it,
               mles_thetas=ths, mles_crlbs=crs, mles_logliks=lls, mles_iterations=its,
               lq_thetas=lq)
    for c in locs.columns:
        out[f"locs_{c}"] = locs[c].to_numpy()
    for em in (False, True):
        lqlocs = glq.locs_from_fits(ids, lq, 7, em)
        for c in lqlocs.columns:
            out[f"lqlocs_em{int(em)}_{c}"] = lqlocs[c].to_numpy()
    gp = np.stack([lq[:, 2], lq[:, 0] + 3, lq[:, 1] + 3, lq[:, 4], lq[:, 5], lq[:, 3]], 1)
    gplocs = glq.locs_from_fits_gpufit(ids, gp, 7, False)
    for c in gplocs.columns:
        out[f"gplocs_{c}"] = gplocs[c].to_numpy()
    ids_nid = ids.copy(); ids_nid["n_id"] = np.arange(len(ids))[::-1].astype(np.int64)
    nlocs = gm.locs_from_fits(ids_nid, th, cr, ll, it, 7)
    out["nidlocs_n_id"] = nlocs["n_id"].to_numpy(); out["nidlocs_x"] = nlocs["x"].to_numpy()
    for bm in (None, "gaussian", "gaussian_iso"):
        n, img = rnd.render(locs, info, oversampling=20, blur_method=bm)
        out[f"render_{bm}_n"] = np.array(n); out[f"render_{bm}_image"] = img
    save("testdata.npz", **out)
    print("testdata: n ids", len(ids), "sum ng", float(ids["net_gradient"].sum()),
          "mean theta", th.mean(0))

def gen_lq(ref):
    """gausslq.fit_spots golden vectors (scipy.optimize.leastsq / MINPACK lmdif path,
    reference gausslq.py:206-289)."""
    from picasso_b200 import testing

    glq = ref["gausslq"]
    out = {}
    for box, n in ((7, 3000), (5, 300), (9, 300), (11, 300), (13, 300)):
        spots = testing.synthetic_spots(n, box, seed=200 + box)
        out[f"b{box}_spots_u16"] = spots.astype(np.uint16)
        out[f"b{box}_thetas"] = glq.fit_spots(spots)
    # noiseless point-sampled float spots (tests/conftest.py:121-154 regime)
    g = np.load(os.path.join(GOLD, "mle_float_spots.npz"))
    out["float_spots"] = g["spots"]
    out["float_thetas"] = glq.fit_spots(g["spots"])
    # photon-converted movie ROIs (non-integer, negative values possible)
    idg = np.load(os.path.join(GOLD, "identify.npz"))
    out["movie_spots"] = idg["mov_b7_spots"]
    out["movie_thetas"] = glq.fit_spots(idg["mov_b7_spots"])
    save("lq.npz", **out)

def gen_render(ref):
    """render.render golden images (reference render.py:37-174, 177-232, 451-575,
    798-853, 1020-1216), unrotated None / gaussian / gaussian_iso."""
    import pandas as pd

    rnd = ref["render"]
    rng = np.random.default_rng(11)
    n = 4000
    locs = pd.DataFrame({
        "frame": rng.integers(0, 100, n).astype(np.uint32),
        "x": rng.uniform(-1, 33, n).astype(np.float32),      # some outside the FOV
        "y": rng.uniform(-1, 25, n).astype(np.float32),
        "lpx": rng.uniform(0.02, 0.4, n).astype(np.float32),
        "lpy": rng.uniform(0.02, 0.4, n).astype(np.float32),
    })
    info = [{"Height": 24, "Width": 32, "Frames": 100, "Pixelsize": 130}]
    out = {c: locs[c].to_numpy() for c in locs.columns}
    cases = {
        "full_os8": dict(oversampling=8),
        "view_os5": dict(oversampling=5, viewport=((4.5, 3.25), (20.125, 30.75))),
        "os1_mbw1": dict(oversampling=1, min_blur_width=1),       # what postprocess.segment uses
        "os2p5_mbw": dict(oversampling=2.5, min_blur_width=0.1),
    }
    for tag, kw in cases.items():
        for bm in (None, "gaussian", "gaussian_iso"):
            k, img = rnd.render(locs, info, blur_method=bm, **kw)
            out[f"{tag}_{bm}_n"] = np.array(k)
            out[f"{tag}_{bm}_image"] = img
    save("render.npz", **out)

def gen_undrift(ref):
    """RCC undrift golden vectors (reference postprocess.py:2824-2961,
    imageprocess.py:27-217, lib.py:2034-2078)."""
    from picasso_b200 import testing

    post, imp = ref["postprocess"], ref["imageprocess"]
    locs, info, _ = testing.synthetic_drift_locs(800, 64, 72, n_clusters=30, locs_per_frame=6,
                                                 seed=3)
    out = {c: locs[c].to_numpy() for c in locs.columns}
    out["info_hwf"] = np.array([64, 72, 800])
    bounds, segments = post.segment(locs, info, 100,
                                    {"blur_method": "gaussian", "min_blur_width": 1}, lambda i: None)
    out["bounds"] = bounds
    out["segments"] = segments.astype(np.float32)     # values are float32 renders
    assert (segments == segments.astype(np.float32)).all()
    n = len(segments)
    sy = np.zeros((n, n)); sx = np.zeros((n, n))
    for i in range(n - 1):
        for j in range(i + 1, n):
            sy[i, j], sx[i, j] = imp.get_image_shift(segments[i], segments[j], 5, 32)
    out["pair_shift_y"] = sy; out["pair_shift_x"] = sx
    shift_y, shift_x = imp.rcc(segments, 32, lambda i: None)
    out["rcc_shift_y"] = shift_y; out["rcc_shift_x"] = shift_x
    out["xcorr_0_1"] = imp.xcorr(segments[0], segments[1])
    out["shift_noroi_0_3"] = np.array(imp.get_image_shift(segments[0], segments[3], 5, None))
    drift, und = post.undrift(locs, info, 100, display=False, segmentation_callback=lambda i: None,
                              rcc_callback=lambda i: None)
    out["drift_x"] = drift["x"].to_numpy(); out["drift_y"] = drift["y"].to_numpy()
    out["undrifted_x"] = und["x"].to_numpy(); out["undrifted_y"] = und["y"].to_numpy()
    save("undrift.npz", **out)


def gen_undrift_c5(ref):
    """Reduced instance of BASELINE config 5: 20 segments x 1024^2, the config-5
    cluster / drift generator scaled to 2 000 frames.  Stores the localizations, the reference's
    per-pair shifts (get_image_shift on the reference's own float64 segment renders), the segment
    shifts of rcc and the final drift of postprocess.undrift -- not the 20 x 1024^2 images."""
    from picasso_b200 import testing

    post, imp = ref["postprocess"], ref["imageprocess"]
    locs, info, _ = testing.synthetic_drift_locs(2000, 1024, 1024, n_clusters=400, locs_per_frame=40.0,
                                                 seed=3, jitter=0.05, lp=0.05)
    out = {c: locs[c].to_numpy() for c in locs.columns}
    out["info_hwf"] = np.array([1024, 1024, 2000])
    bounds, segments = post.segment(locs, info, 100,
                                    {"blur_method": "gaussian", "min_blur_width": 1}, lambda i: None)
    out["bounds"] = bounds
    n = len(segments)
    sy = np.zeros((n, n)); sx = np.zeros((n, n))
    for i in range(n - 1):
        for j in range(i + 1, n):
            sy[i, j], sx[i, j] = imp.get_image_shift(segments[i], segments[j], 5, 32)
    out["pair_shift_y"] = sy; out["pair_shift_x"] = sx
    shift_y, shift_x = imp.rcc(segments, 32, lambda i: None)
    out["rcc_shift_y"] = shift_y; out["rcc_shift_x"] = shift_x
    drift, und = post.undrift(locs, info, 100, display=False, segmentation_callback=lambda i: None,
                              rcc_callback=lambda i: None)
    out["drift_x"] = drift["x"].to_numpy(); out["drift_y"] = drift["y"].to_numpy()
    out["segment_sums"] = segments.sum((1, 2))
    save("undrift_c5.npz", **out)
    print("undrift_c5:", len(locs), "locs", n, "segments")


def zfit_problem(n=3000, seed=11):
    """Synthetic astigmatism calibration + 2-D fitted localizations (shared with the tests)."""
    from picasso_b200 import testing

    return testing.synthetic_zfit_locs(n, seed)


def gen_zfit(ref):
    import pandas as pd

    zf = ref["zfit"]
    locs, info, calib = zfit_problem()
    out = {k: locs[k].to_numpy() for k in locs.columns}
    out["cx"] = np.array(calib["X Coefficients"]); out["cy"] = np.array(calib["Y Coefficients"])
    out["magnification"] = np.float64(calib["Magnification factor"])
    for tag, method, drop_unc, flt in (("lq_f0", "gausslq", True, 0), ("lq_f2", "gausslq", True, 2),
                                       ("mle_f0", "gaussmle", True, 0), ("mleunc_f2", "gaussmle", False, 2)):
        l = locs.drop(columns=["sx_unc", "sy_unc"]) if drop_unc else locs
        res = zf._fit_z(l, info, calib, calib["Magnification factor"], 130, fitting_method=method, filter=flt)
        out[f"{tag}_index"] = res.index.to_numpy()
        for c in ("z", "d_zcalib", "lpz"):
            out[f"{tag}_{c}"] = res[c].to_numpy()
        print(tag, len(res), {c: res[c].dtype for c in ("z", "d_zcalib", "lpz")})
    # raw minimiser outputs (before ensure_sanity / filter)
    from scipy.optimize import minimize_scalar
    sx = locs["sx"].to_numpy(); sy = locs["sy"].to_numpy()
    cx, cy = out["cx"], out["cy"]
    zr = np.zeros(len(locs)); fr = np.zeros(len(locs)); nf = np.zeros(len(locs), np.int32)
    for i in range(len(locs)):
        r = minimize_scalar(zf._fit_z_target, bounds=[-1000, 1000], args=(sx[i], sy[i], cx, cy))
        zr[i], fr[i], nf[i] = r.x, r.fun, r.nfev
    out["raw_z"] = zr; out["raw_fun"] = fr; out["raw_nfev"] = nf
    save("zfit.npz", **out)


def gen_aim(ref):
    """aim.aim on 2-D and 3-D localizations, plus the per-segment intersection counts of every
    round (recorded by wrapping the reference's _point_intersect_* helpers)."""
    from picasso_b200 import testing

    aim = ref["aim"]
    out = {}
    for tag, with_z in (("2d", False), ("3d", True)):
        locs, info, _ = testing.synthetic_aim_locs(with_z=with_z)
        rec2, rec3 = [], []
        o2, o3 = aim._point_intersect_2d, aim._point_intersect_3d

        def w2(*a, **k):
            r = o2(*a, **k); rec2.append(np.array(r)); return r

        def w3(*a, **k):
            r = o3(*a, **k); rec3.append(np.array(r)); return r

        aim._point_intersect_2d, aim._point_intersect_3d = w2, w3
        try:
            und, new_info, drift = aim.aim(locs, info, segmentation=100)
        finally:
            aim._point_intersect_2d, aim._point_intersect_3d = o2, o3
        out[f"{tag}_roi_cc"] = np.stack(rec2)
        if rec3:
            out[f"{tag}_roi_cc_z"] = np.stack(rec3)
        for c in drift.columns:
            out[f"{tag}_drift_{c}"] = drift[c].to_numpy()
        for c in ("x", "y") + (("z",) if with_z else ()):
            out[f"{tag}_und_{c}"] = und[c].to_numpy()
        print(tag, len(locs), out[f"{tag}_roi_cc"].shape, {c: und[c].dtype for c in und.columns},
              drift.dtypes.to_dict(), new_info[-1])
    # a second geometry: non-default intersect_d / roi_r, unsorted frames, frame offset
    locs, info, _ = testing.synthetic_aim_locs(n_frames=900, Y=48, X=80, n_clusters=30, seed=9)
    rng = np.random.default_rng(1)
    locs = locs.iloc[rng.permutation(len(locs))].reset_index(drop=True)
    locs["frame"] += 7
    und, new_info, drift = aim.aim(locs, info, segmentation=150, intersect_d=0.2, roi_r=0.55)
    out["alt_drift_x"] = drift["x"].to_numpy(); out["alt_drift_y"] = drift["y"].to_numpy()
    out["alt_und_x"] = und["x"].to_numpy(); out["alt_und_y"] = und["y"].to_numpy()
    out["alt_perm_frame"] = locs["frame"].to_numpy()
    out["alt_perm_x"] = locs["x"].to_numpy(); out["alt_perm_y"] = locs["y"].to_numpy()
    save("aim.npz", **out)


def gen_link(ref):
    """postprocess.link (average mode) and the raw link groups on blinking-site data."""
    from picasso_b200 import testing

    post = ref["postprocess"]
    out = {}
    for tag, kw in (("plain", {}), ("group", {"with_group": True}), ("f64", {"f64_xy": True, "seed": 8})):
        locs, info = testing.synthetic_link_locs(**kw)
        sl = locs.sort_values(kind="quicksort", by="frame")
        group = sl["group"].to_numpy() if "group" in sl.columns else np.zeros(len(sl), dtype=np.int32)
        for dark in (3, 1):
            lg = post._get_link_groups(sl["frame"].to_numpy(), sl["x"].to_numpy(), sl["y"].to_numpy(), 0.05, dark, group)
            out[f"{tag}_lg_dark{dark}"] = lg
        out[f"{tag}_sorted_index"] = sl.index.to_numpy()
        linked = post.link(locs, info, r_max=0.05, max_dark_time=3)
        out[f"{tag}_linked_index"] = linked.index.to_numpy()
        for c in linked.columns:
            out[f"{tag}_linked_{c}"] = linked[c].to_numpy()
        print(tag, len(locs), "->", len(linked), {c: str(linked[c].dtype) for c in linked.columns})
    save("link.npz", **out)


GENERATORS = {"link": gen_link, "aim": gen_aim, "zfit": gen_zfit, "mle": gen_mle, "identify": gen_identify, "testdata": gen_testdata, "lq": gen_lq, "render": gen_render, "undrift": gen_undrift, "undrift_c5": gen_undrift_c5}


def main():
    which = sys.argv[1:] or list(GENERATORS)
    ref = ref_import.import_reference()
    for w in which:
        GENERATORS[w](ref)


if __name__ == "__main__":
    main()
