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
