"""Least-squares 2-D Gaussian spot fitting on B200.

Drop-in for ``picasso.gausslq`` (reference picasso/gausslq.py): ``fit_spot`` :206,
``fit_spots`` :247, ``fit_spots_parallel`` :292, ``fit_spots_gpufit`` :346,
``fits_from_futures`` :398, ``locs_from_fits`` :404, ``locs_from_fits_gpufit`` :487,
``localization_precision`` :547, ``sigma_uncertainty`` :592.  The optimiser the
reference calls (scipy.optimize.leastsq == MINPACK lmdif, ftol = xtol = 1e-2) runs
per spot inside the CUDA kernel csrc/lq_fit.cu.
"""
from __future__ import annotations

import ctypes as C
from concurrent import futures as _futures
from typing import Callable, Literal

import numpy as np
import pandas as pd

from . import _lib

# The reference probes for the Windows Gpufit DLL (gausslq.py:25-30).  Here the
# GPU path is native, so the "gpufit" entry points are always available.
GPUFIT_INSTALLED = True


def _declare(lib):
    if getattr(lib, "_lq_declared", False):
        return
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.pb_lq_fit.argtypes = [sz, i32, vp, vp, vp, vp]
    lib.pb_lq_fit.restype = i32
    lib.pb_lq_fit_dev.argtypes = [sz, i32, vp, vp, vp, vp, vp]
    lib.pb_lq_fit_dev.restype = i32
    lib._lq_declared = True


def _fit(spots, want_info=False):
    lib = _lib.load()
    _declare(lib)
    _lib.require_gpu()
    spots = np.ascontiguousarray(spots, dtype=np.float32)
    if spots.ndim != 3 or spots.shape[1] != spots.shape[2]:
        raise ValueError("spots must have shape (n_spots, size, size)")
    n, box, _ = spots.shape
    # the reference pre-fills with NaN (gausslq.py:277-278); here the kernel writes every row, and
    # the array is backed by pooled page-locked memory so the download is a direct DMA
    theta = _lib.pinned_empty((n, 6), np.float32)
    infos = np.zeros(n, np.int32) if want_info else None
    nfevs = np.zeros(n, np.int32) if want_info else None
    if n:
        _lib.check(lib.pb_lq_fit(n, box, _lib.ptr(spots), _lib.ptr(theta),
                                 _lib.ptr(infos) if want_info else None,
                                 _lib.ptr(nfevs) if want_info else None))
    if want_info:
        return theta, infos, nfevs
    return theta


def fit_spot(spot) -> np.ndarray:
    """Fit one spot; returns ``[x, y, photons, bg, sx, sy]`` (float64 like the
    reference's ``leastsq`` result, gausslq.py:206-244) with x, y relative to the
    box centre."""
    spot = np.asarray(spot)
    return _fit(spot[None])[0].astype(np.float64)


def fit_spots(spots, progress_callback: Callable[[int], None] | Literal["console"] | None = None):
    """Fit all spots (reference gausslq.py:247-289): float32 ``(n, 6)`` array with
    columns ``[x, y, photons, bg, sx, sy]``."""
    theta = _fit(spots)
    n = len(theta)
    if progress_callback == "console":
        from tqdm import tqdm

        with tqdm(total=n, desc="Fitting...", unit="spot") as bar:
            bar.update(n)
    elif callable(progress_callback):
        for i in range(n):
            progress_callback(i)
    return theta


def fit_spots_parallel(spots, asynch: bool = False):
    """Reference gausslq.py:292-343 fans the fit out over a process pool.  Here
    one GPU call does the whole batch; with ``asynch=True`` it runs on a worker
    thread and a one-element list of futures is returned (``fits_from_futures``
    stacks the results as in the reference)."""
    if asynch:
        ex = _futures.ThreadPoolExecutor(1)
        fs = [ex.submit(_lib.on_callers_device(fit_spots), spots)]
        ex.shutdown(wait=False)
        return fs
    return fit_spots(spots)


def fits_from_futures(futures) -> np.ndarray:
    """Collect results from futures and stack them (reference gausslq.py:398-401)."""
    return np.vstack([f.result() for f in futures])


def _initial_parameters_gpufit(spots, size: int) -> np.ndarray:
    """Start values of the reference's Gpufit path (gausslq.py:128-148), kept for
    API compatibility: [photons(amplitude), x, y, sx, sy, bg]."""
    center = (size / 2.0) - 0.5
    width = np.amax([size / 5.0, 1.0])
    p = np.empty((len(spots), 6), dtype=np.float32)
    smax = np.amax(spots, axis=(1, 2))
    smin = np.amin(spots, axis=(1, 2))
    p[:, 0] = smax - smin
    p[:, 1] = center
    p[:, 2] = center
    p[:, 3] = width
    p[:, 4] = width
    p[:, 5] = smin
    return p


def fit_spots_gpufit(spots, return_info: bool = False) -> np.ndarray:
    """The reference's Gpufit path (gausslq.py:346-395): start values
    ``_initial_parameters_gpufit``, Gpufit's float32 Levenberg-Marquardt on ``GAUSS_2D_ELLIPTIC`` with
    tolerance 1e-2 and at most 20 iterations, then amplitude -> photons (``* 2 pi sx sy``).  Returns
    ``[photons, x, y, sx, sy, bg]`` with x, y in box coordinates (0 .. size-1), so
    ``locs_from_fits_gpufit`` applies unchanged.

    The vendored Gpufit 1.2.0 binary (Windows DLL, no source) cannot run on Linux / B200; the kernel
    (csrc/gpufit_lm.cu) runs Gpufit's PUBLISHED algorithm -- its own start values and trajectory, not
    the MINPACK result relabelled.  Parity with the binary itself is unpinned;
    the CPU oracle restates the same algorithm, and both agree with ``fit_spots`` within the LQ
    tolerance (tests/test_lq_gpu.py)."""
    lib = _lib.load()
    _lib.require_gpu()
    vp, i32, sz, f32 = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    lib.pb_gpufit_fit.argtypes = [sz, i32, vp, f32, i32, vp, vp, vp, vp]
    lib.pb_gpufit_fit.restype = i32
    spots = np.ascontiguousarray(spots, dtype=np.float32)
    if spots.ndim != 3 or spots.shape[1] != spots.shape[2]:
        raise ValueError("spots must have shape (n_spots, size, size)")
    n, box, _ = spots.shape
    params = _lib.pinned_empty((n, 6), np.float32)
    states = np.zeros(n, np.int32) if return_info else None
    chi2 = np.zeros(n, np.float32) if return_info else None
    nit = np.zeros(n, np.int32) if return_info else None
    if n:
        p = lambda a: _lib.ptr(a) if a is not None else None      # noqa: E731
        _lib.check(lib.pb_gpufit_fit(n, box, _lib.ptr(spots), 1e-2, 20, _lib.ptr(params), p(states), p(chi2),
                                     p(nit)))
    if return_info:
        return params, states, chi2, nit
    return params


def localization_precision(photons, s, s_orth, bg, em: bool):
    """Mortensen et al. (2010) precision of an unweighted 2-D Gaussian fit with a
    diagonal covariance (reference gausslq.py:547-589); EM gain doubles the variance."""
    sa2 = s ** 2 + 1 / 12
    sa = sa2 ** 0.5
    sa_orth = (s_orth ** 2 + 1 / 12) ** 0.5
    v = sa2 * (16 / 9 + (8 * np.pi * sa * sa_orth * bg) / photons) / photons
    if em:
        v *= 2
    with np.errstate(invalid="ignore"):
        return np.sqrt(v)


def sigma_uncertainty(sigma, sigma_orth, photons, bg):
    """Standard error of the fitted sigma for the LQ model (reference gausslq.py:592-633)."""
    sa2 = sigma ** 2 + 1 / 12
    sa = sa2 ** 0.5
    sa_orth = (sigma_orth ** 2 + 1 / 12) ** 0.5
    var_sa2 = sa2 ** 2 / photons * (512 / 81 + (64 * np.pi * sa * sa_orth * bg) / (3 * photons))
    return np.sqrt(var_sa2 / (4 * sigma ** 2))


def _locs_table(identifications, x, y, photons, sx, sy, bg, em):
    f32 = np.float32
    lpx = localization_precision(photons, sx, sy, bg, em=em)
    lpy = localization_precision(photons, sy, sx, bg, em=em)
    big = np.maximum(sx, sy)
    small = np.minimum(sx, sy)
    cols = {
        "frame": identifications["frame"].astype(np.uint32),
        "x": x.astype(f32),
        "y": y.astype(f32),
        "photons": photons.astype(f32),
        "sx": sx.astype(f32),
        "sy": sy.astype(f32),
        "bg": bg.astype(f32),
        "lpx": lpx.astype(f32),
        "lpy": lpy.astype(f32),
        "ellipticity": ((big - small) / big).astype(f32),
        "net_gradient": identifications["net_gradient"].astype(f32),
    }
    return cols


def locs_from_fits(identifications: pd.DataFrame, theta, box: int, em: bool) -> pd.DataFrame:
    """Localization table from LQ fits (reference gausslq.py:404-484): x, y are the
    fitted offsets plus the identification pixel (no box offset); sorted by
    ``n_id`` when present, else ``frame``."""
    cols = _locs_table(identifications, theta[:, 0] + identifications["x"],
                       theta[:, 1] + identifications["y"], theta[:, 2], theta[:, 4], theta[:, 5],
                       theta[:, 3], em)
    if "n_id" in identifications.columns:
        cols["n_id"] = identifications["n_id"].astype(np.uint32)
        locs = pd.DataFrame(cols)
        locs.sort_values(by="n_id", kind="quicksort", inplace=True)
    else:
        locs = pd.DataFrame(cols)
        locs.sort_values(by="frame", kind="quicksort", inplace=True)
    return locs


def locs_from_fits_gpufit(identifications: pd.DataFrame, theta, box: int, em: bool) -> pd.DataFrame:
    """Localization table from the Gpufit column layout ``[photons, x, y, sx, sy, bg]``
    (reference gausslq.py:487-544): x, y minus ``box // 2``; sorted by frame."""
    half = int(box / 2)
    cols = _locs_table(identifications, theta[:, 1] + identifications["x"] - half,
                       theta[:, 2] + identifications["y"] - half, theta[:, 0], theta[:, 3],
                       theta[:, 4], theta[:, 5], em)
    locs = pd.DataFrame(cols)
    locs.sort_values(by="frame", kind="quicksort", inplace=True)
    return locs
