"""RCC drift correction on B200.

Drop-in for the hot-path part of ``picasso.postprocess`` (reference
picasso/postprocess.py): ``n_segments`` :2824, ``segment`` :2846, ``undrift`` :2903,
``apply_drift`` :3171.  Segment images are rendered by the CUDA renderer, the pairwise
cross-correlations by cuFFT (picasso_b200.imageprocess.rcc); the cubic spline and the
subtraction are host numpy/scipy exactly as in the reference.
"""
from __future__ import annotations

import warnings
from typing import Callable

import numpy as np
import pandas as pd

from . import imageprocess, lib, render


def n_segments(info, segmentation: int) -> int:
    """Number of temporal segments: ``int(np.round(Frames / segmentation))``
    (banker's rounding; reference postprocess.py:2824-2843)."""
    n_frames = lib.get_from_metadata(info, "Frames")
    return int(np.round(n_frames / segmentation))


def segment(locs: pd.DataFrame, info, segmentation: int, kwargs: dict = {},
            callback: Callable[[int], None] | None = None):
    """Split ``locs`` into temporal segments and render each one (reference
    postprocess.py:2846-2900).  Returns ``(bounds uint32 (n_seg+1,), segments float64
    (n_seg, Y, X))``; segment i holds frames ``bounds[i] <= frame < bounds[i+1]`` (so the
    very last frame is dropped, as in the reference); ``callback`` gets 0..n_seg."""
    Y = info[0]["Height"]
    X = info[0]["Width"]
    n_frames = info[0]["Frames"]
    n_seg = n_segments(info, segmentation)
    bounds = np.linspace(0, n_frames - 1, n_seg + 1, dtype=np.uint32)
    segments = np.zeros((n_seg, Y, X))
    bar = None
    if callback is None:
        from tqdm import trange

        it = trange(n_seg, desc="Generating segments", unit="segments")
    else:
        callback(0)
        it = range(n_seg)
    frames = locs["frame"].to_numpy()
    for i in it:
        sel = (frames >= bounds[i]) & (frames < bounds[i + 1])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            _, segments[i] = render.render(locs[sel], info, **kwargs)
        if callback is not None:
            callback(i + 1)
    return bounds, segments


def _apply_drift(locs: pd.DataFrame, drift: pd.DataFrame) -> pd.DataFrame:
    """``locs[c] -= drift[c].iloc[frames]`` (reference postprocess.py:3159-3168); the per-frame
    values are gathered with one numpy take (same values; pandas' positional indexer costs 5x
    more on millions of rows).  Negative frames index from the end like ``.iloc``."""
    frames = locs["frame"].to_numpy().astype(np.intp)
    n = len(drift)
    if len(frames) and (frames.min() < -n or frames.max() >= n):
        raise IndexError("positional indexers are out-of-bounds")
    for c in ("x", "y", "z"):
        if c in drift.columns and (c != "z" or c in locs.columns):
            locs[c] -= drift[c].to_numpy()[frames]
    return locs


def apply_drift(locs: pd.DataFrame, info, *, drift):
    """Subtract a per-frame drift from the localizations (reference
    postprocess.py:3171-3218): ``drift`` is a DataFrame with columns x, y[, z] or an
    array ``(n_frames, 2|3)``."""
    assert isinstance(drift, (pd.DataFrame, np.ndarray)), "Drift must be a DataFrame or numpy array"
    n_frames = lib.get_from_metadata(info, "Frames", raise_error=True)
    if isinstance(drift, pd.DataFrame):
        required = {"x", "y"}
        if not required.issubset(drift.columns):
            raise ValueError(f"Drift DataFrame must contain columns {required}")
    else:
        if not (drift.shape[1] in [2, 3] and drift.shape[0] == n_frames):
            raise ValueError("Drift array must have shape (n_frames, 2) for x and y drift, "
                             "or (n_frames, 3) for x, y, and z drift.")
        drift = pd.DataFrame(drift, columns=["x", "y"] + (["z"] if drift.shape[1] == 3 else []))
    return _apply_drift(locs, drift)


def undrift(locs: pd.DataFrame, info, segmentation: int, display: bool = True,
            segmentation_callback: Callable[[int], None] | None = None,
            rcc_callback: Callable[[int], None] | None = None, _shifts_fn=None):
    """Undrift by RCC (reference postprocess.py:2903-2961): render segments
    (gaussian blur, min_blur_width=1, oversampling 1), cross-correlate all pairs
    (max_shift 32), spline the segment shifts over all frames and subtract.  Returns
    ``(drift DataFrame{x, y} of length Frames, undrifted locs copy)``.  ``display`` is
    accepted for compatibility; plotting is GUI code outside the hot path."""
    from scipy import interpolate

    locs = locs.copy()
    # same work as segment(..., blur_method="gaussian", min_blur_width=1) followed by
    # imageprocess.rcc(segments, 32), but the segment images never leave the GPU
    n_frames = info[0]["Frames"]
    n_seg = n_segments(info, segmentation)
    bounds = np.linspace(0, n_frames - 1, n_seg + 1, dtype=np.uint32)
    if segmentation_callback is not None:
        for i in range(n_seg + 1):
            segmentation_callback(i)
    if rcc_callback is None:
        rcc_callback = lambda _i: None
    # _shifts_fn: multi-GPU callers (picasso_b200.distributed.undrift_sharded) shard the pairs
    rcc_of_locs = _shifts_fn if _shifts_fn is not None else imageprocess._rcc_of_locs
    shift_y, shift_x = rcc_of_locs(locs, info, bounds, 1, 32, rcc_callback)
    t = (bounds[1:] + bounds[:-1]) / 2
    drift_x_pol = interpolate.InterpolatedUnivariateSpline(t, shift_x, k=3)
    drift_y_pol = interpolate.InterpolatedUnivariateSpline(t, shift_y, k=3)
    t_inter = np.arange(info[0]["Frames"])
    drift = pd.DataFrame({"x": drift_x_pol(t_inter), "y": drift_y_pol(t_inter)})
    locs = apply_drift(locs, info, drift=drift)
    return drift, locs


# ---- linking (reference postprocess.py:2007-2821) ------------------------------------------
_LINK_DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.uint32): 2,
                np.dtype(np.int32): 3}


def _declare_link(l):
    if getattr(l, "_link_declared", False):
        return
    import ctypes as C

    vp, i32 = C.c_void_p, C.c_int
    l.pb_link_groups.argtypes = [C.c_size_t, vp, vp, vp, i32, vp, C.c_double, C.c_longlong, vp, C.POINTER(i32)]
    l.pb_link_groups.restype = i32
    l.pb_link_reduce.argtypes = [C.c_size_t, vp, i32, i32, vp, vp, vp, vp]
    l.pb_link_reduce.restype = i32
    l._link_declared = True


def _get_link_groups(frame, x, y, d_max: float, max_dark_time: int, group):
    """Link group of every localization (reference ``_get_link_groups``, postprocess.py:2440-2507);
    ``frame`` must be sorted ascending.  Runs on the GPU (csrc/link.cu), bit-exact."""
    import ctypes as C

    from . import _lib

    l = _lib.load()
    _declare_link(l)
    _lib.require_gpu()
    frame = np.ascontiguousarray(frame, dtype=np.int64)
    x = np.asarray(x)
    f64 = int(x.dtype == np.float64 or np.asarray(y).dtype == np.float64)
    dt = np.float64 if f64 else np.float32
    x = np.ascontiguousarray(x, dtype=dt)
    y = np.ascontiguousarray(y, dtype=dt)
    group = np.ascontiguousarray(group, dtype=np.int32)
    link_group = np.empty(len(frame), np.int32)
    n_groups = C.c_int(0)
    _lib.check(l.pb_link_groups(len(frame), _lib.ptr(frame), _lib.ptr(x), _lib.ptr(y), f64, _lib.ptr(group),
                                float(d_max), int(max_dark_time), _lib.ptr(link_group), C.byref(n_groups)))
    return link_group


get_link_groups = _get_link_groups


def _group_reduce(link_group, n_groups, columns, ops):
    """Per-group reductions of several columns in one GPU pass (``pb_link_reduce``)."""
    import ctypes as C

    from . import _lib

    l = _lib.load()
    _declare_link(l)
    cols, orig = [], []
    for c in columns:
        c = np.ascontiguousarray(c)
        orig.append(c.dtype)
        if c.dtype not in _LINK_DTYPES:       # e.g. int64 frames / groups: values fit 32 bits
            c = c.astype(np.float64 if c.dtype.kind == "f" else (np.uint32 if c.dtype.kind == "u" else np.int32))
        cols.append(c)
    outs = [np.zeros(n_groups, dtype=c.dtype) for c in cols]
    k = len(cols)
    if k and n_groups:
        VP = C.c_void_p * k
        I = C.c_int * k
        _lib.check(l.pb_link_reduce(len(link_group), _lib.ptr(link_group), int(n_groups), k,
                                    VP(*[c.ctypes.data for c in cols]), I(*[_LINK_DTYPES[c.dtype] for c in cols]),
                                    I(*ops), VP(*[o.ctypes.data for o in outs])))
    return [o if o.dtype == d else o.astype(d) for o, d in zip(outs, orig)]


def _link_loc_groups(locs: pd.DataFrame, info, link_group, remove_ambiguous_lengths: bool = True):
    """Combine linked localizations into binding events (reference ``_link_loc_groups``,
    postprocess.py:2680-2821): weighted mean position, summed photons / bg, mean widths, ``len``,
    ``n``, ``photon_rate``.  The per-group accumulations run on the GPU in the reference's order
    and dtypes; the elementwise steps are the reference's numpy expressions."""
    from collections import OrderedDict

    link_group = np.ascontiguousarray(link_group, dtype=np.int32)
    n_groups = int(link_group.max()) + 1
    SUM, MIN, MAX, LAST = 0, 1, 2, 3
    jobs = OrderedDict()          # name -> (column, op)
    jobs["n"] = (np.ones(len(link_group), np.uint32), SUM)
    has = lambda c: c in locs.columns   # noqa: E731
    col = lambda c: locs[c].to_numpy()  # noqa: E731
    if has("frame"):
        jobs["first"] = (col("frame"), MIN)
        jobs["last"] = (col("frame"), MAX)
    weights = {}
    for c, lp in (("x", "lpx"), ("y", "lpy")):
        if has(c):
            weights[c] = 1 / col(lp) ** 2
            jobs["w_" + c] = (weights[c], SUM)
            jobs["wsum_" + c] = (col(c) * weights[c], SUM)
    if has("z") and has("lpz"):
        weights["z"] = 1 / col("lpz") ** 2
        jobs["w_z"] = (weights["z"], SUM)
        jobs["wsum_z"] = (col("z") * weights["z"], SUM)
    plain_sum = [c for c in ("photons", "bg") if has(c)]
    plain_mean = [c for c in ("sx", "sy", "ellipticity", "net_gradient", "likelihood", "iterations", "d_zcalib")
                  if has(c)]
    if has("z") and not has("lpz"):
        plain_mean.append("z")
    for c in plain_sum + plain_mean:
        jobs["sum_" + c] = (col(c), SUM)
    if has("group"):
        jobs["group"] = (col("group"), LAST)
    res = dict(zip(jobs, _group_reduce(link_group, n_groups, [v[0] for v in jobs.values()],
                                       [v[1] for v in jobs.values()])))
    n_ = res["n"]

    def f32_div(a, b):
        out = np.empty(n_groups, dtype=np.float32)      # "this ensures float32 after the division"
        out[:] = a / b
        return out

    columns = OrderedDict()
    if has("frame"):
        columns["frame"] = res["first"]
    for c in ("x", "y"):
        if has(c):
            columns[c] = f32_div(res["wsum_" + c], res["w_" + c])
    if has("photons"):
        columns["photons"] = res["sum_photons"]
    for c in ("sx", "sy"):
        if has(c):
            columns[c] = f32_div(res["sum_" + c], n_)
    if has("bg"):
        columns["bg"] = res["sum_bg"]
    if has("x"):
        columns["lpx"] = np.sqrt(1 / res["w_x"])
    if has("y"):
        columns["lpy"] = np.sqrt(1 / res["w_y"])
    for c in ("ellipticity", "net_gradient", "likelihood", "iterations"):
        if has(c):
            columns[c] = f32_div(res["sum_" + c], n_)
    if has("z"):
        if has("lpz"):
            columns["z"] = f32_div(res["wsum_z"], res["w_z"])
            columns["lpz"] = np.sqrt(1 / res["w_z"])
        else:
            columns["z"] = f32_div(res["sum_z"], n_)
    if has("d_zcalib"):
        columns["d_zcalib"] = f32_div(res["sum_d_zcalib"], n_)
    if has("group"):
        columns["group"] = res["group"]
    if has("frame"):
        columns["len"] = res["last"] - res["first"] + 1
    columns["n"] = n_
    if has("photons"):
        columns["photon_rate"] = np.float32(columns["photons"] / n_)
    linked_locs = pd.DataFrame(columns)
    if remove_ambiguous_lengths:
        valid = np.logical_and(res["first"] > 0, res["last"] < info[0]["Frames"])
        linked_locs = linked_locs[valid]
    return linked_locs


link_loc_groups = _link_loc_groups


def link(locs: pd.DataFrame, info, r_max: float = 0.05, max_dark_time: int = 3,
         combine_mode: str = "average", remove_ambiguous_lengths: bool = True) -> pd.DataFrame:
    """Link localizations into binding events (reference ``link``, postprocess.py:2007-2072)."""
    if len(locs) == 0:
        linked_locs = locs.copy()
        if "frame" in locs.columns:
            linked_locs["len"] = np.array([], dtype=np.int32)
            linked_locs["n"] = np.array([], dtype=np.int32)
        if "photons" in locs.columns:
            linked_locs["photon_rate"] = np.array([], dtype=np.float32)
        return linked_locs
    locs = locs.sort_values(kind="quicksort", by="frame")
    group = locs["group"].to_numpy() if "group" in locs.columns else np.zeros(len(locs), dtype=np.int32)
    link_group = _get_link_groups(locs["frame"].to_numpy(), locs["x"].to_numpy(), locs["y"].to_numpy(),
                                  r_max, max_dark_time, group)
    if combine_mode == "average":
        return _link_loc_groups(locs, info, link_group, remove_ambiguous_lengths=remove_ambiguous_lengths)
    if combine_mode == "refit":
        raise NotImplementedError("Refit mode is not implemented yet. Please use 'average' mode.")
    return None
