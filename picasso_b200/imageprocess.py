"""FFT cross-correlation and RCC on B200.

Drop-in for the hot-path part of ``picasso.imageprocess`` (reference
picasso/imageprocess.py): ``xcorr`` :27, ``get_image_shift`` :53, ``rcc`` :160.  The
transforms run in cuFFT (one forward transform per segment, one inverse per pair,
csrc/rcc.cu); the 5-parameter peak fit of the 5x5 window stays on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

import numpy as np

from . import _lib, lib


def _declare(l):
    if getattr(l, "_rcc_declared", False):
        return
    vp, i32 = C.c_void_p, C.c_int
    l.pb_rcc_windows.argtypes = [i32, i32, i32, vp, i32, i32, i32, i32, vp, vp]
    l.pb_rcc_windows.restype = i32
    l.pb_undrift_windows_pairs.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, C.c_double, i32, i32, i32, i32,
                                           i32, vp, vp, vp, vp, vp]
    l.pb_undrift_windows_pairs.restype = i32
    l.pb_undrift_peaks_pairs.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, C.c_double, i32, i32, i32, i32,
                                         i32, vp, vp, vp, vp, vp]
    l.pb_undrift_peaks_pairs.restype = i32
    l.pb_undrift_windows.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, C.c_double, i32, i32, i32,
                                     i32, vp, vp, vp]
    l.pb_undrift_windows.restype = i32
    l._rcc_declared = True


def _crop_geometry(Y, X, roi):
    """Rows/cols kept by the reference's centre crop (imageprocess.py:88-101)."""
    Y_ = X_ = 0
    if roi is not None:
        Y_ = int((Y - roi) / 2)
        X_ = int((X - roi) / 2)
        if Y_ <= 0:
            Y_ = 0
        if X_ <= 0:
            X_ = 0
    return Y_, X_, Y - 2 * Y_, X - 2 * X_


def _windows(segments, roi):
    """Cropped, fftshift-ed, normalised correlation windows of all i<j pairs."""
    l = _lib.load()
    _declare(l)
    _lib.require_gpu()
    seg = np.ascontiguousarray(segments, dtype=np.float32)
    n_seg, Y, X = seg.shape
    Y_, X_, H, W = _crop_geometry(Y, X, roi)
    n_pairs = n_seg * (n_seg - 1) // 2
    win = np.zeros((n_pairs, H, W), dtype=np.float32)
    sums = np.zeros(n_seg, dtype=np.float64)
    _lib.check(l.pb_rcc_windows(n_seg, Y, X, _lib.ptr(seg), Y_, X_, H, W, _lib.ptr(win),
                                _lib.ptr(sums)))
    return win, sums, (Y_, X_)


def xcorr(imageA, imageB):
    """Cross-correlation of two images (reference imageprocess.py:27-50):
    ``fftshift(real(ifft2(fft2(A) * conj(fft2(B))))) / sqrt(A.size)``."""
    stack = np.stack([np.asarray(imageA), np.asarray(imageB)])
    win, _, _ = _windows(stack, None)
    return win[0].astype(np.float64)


def _gauss_peak_fit(window):
    """Fit ``a * exp(-0.5 * ((x - xc)^2 + (y - yc)^2) / s^2) + b`` to a square window
    centred on the brightest pixel, start ``[max, 0, 0, 1, min]``, bounds a, s, b >= 0
    -- the model, start values and bounds of the reference's ``curve_fit`` call
    (imageprocess.py:119-135).  Returns (xc, yc)."""
    from scipy.optimize import curve_fit

    k = window.shape[0] // 2
    y, x = np.mgrid[-k:k + 1, -k:k + 1]

    def model(coords, a, xc, yc, s, b):
        xx, yy = coords
        return (a * np.exp(-0.5 * ((xx - xc) ** 2 + (yy - yc) ** 2) / s ** 2) + b).ravel()

    p0 = [window.max(), 0, 0, 1, window.min()]
    bounds = ([0, -np.inf, -np.inf, 0, 0], [np.inf] * 5)
    popt, _ = curve_fit(model, (x, y), window.ravel(), p0=p0, bounds=bounds)
    return popt[1], popt[2]


def _gauss_peak_fit_batch(windows, max_iter: int = 60):
    """Vectorised Levenberg-Marquardt fit of the same model / start values / bounds as
    ``_gauss_peak_fit`` for a stack of (P, k, k) windows at once (RCC fits n(n-1)/2 peaks;
    19 900 scipy ``curve_fit`` calls would cost more than all the FFTs on the GPU).

    Returns ``(xc, yc, ok)``; ``ok[p]`` is False where the batched solver did not converge
    to an interior optimum (a bound active, singular normal equations, no convergence) --
    the caller re-fits those windows with scipy's bounded ``curve_fit`` exactly like the
    reference.  Both solvers minimise the same convex-near-the-peak least-squares problem
    to ~1e-8, so converged fits agree far below the 1e-3 px parity bar (tests/test_peakfit_cpu.py).
    """
    w = np.asarray(windows, dtype=np.float64)
    P, k, _ = w.shape
    h = k // 2
    yy, xx = np.mgrid[-h:h + 1, -h:h + 1]
    xx = xx.ravel()[None, :].astype(np.float64)
    yy = yy.ravel()[None, :].astype(np.float64)
    d = w.reshape(P, k * k)
    p = np.stack([d.max(1), np.zeros(P), np.zeros(P), np.ones(P), d.min(1)], 1)   # a xc yc s b
    lam = np.full(P, 1e-3)

    def model_jac(p):
        a, xc, yc, s, b = (p[:, i:i + 1] for i in range(5))
        dx, dy = xx - xc, yy - yc
        r2 = dx * dx + dy * dy
        E = np.exp(-0.5 * r2 / (s * s))
        m = a * E + b
        J = np.stack([E, a * E * dx / (s * s), a * E * dy / (s * s), a * E * r2 / (s ** 3),
                      np.ones_like(E)], 2)
        return m, J

    m, J = model_jac(p)
    res = d - m
    cost = (res * res).sum(1)
    active = np.ones(P, bool)
    converged = np.zeros(P, bool)
    eye = np.eye(5)[None]
    for _ in range(max_iter):
        if not active.any():
            break
        idx = np.flatnonzero(active)
        Ja, ra = J[idx], res[idx]
        JTJ = np.einsum("pij,pik->pjk", Ja, Ja)
        g = np.einsum("pij,pi->pj", Ja, ra)
        A = JTJ + lam[idx, None, None] * (eye * np.maximum(np.einsum("pjj->pj", JTJ), 1e-300)[:, None, :])
        try:
            step = np.linalg.solve(A, g[:, :, None])[:, :, 0]
        except np.linalg.LinAlgError:
            step = np.zeros_like(g)
            for q in range(len(idx)):
                try:
                    step[q] = np.linalg.solve(A[q], g[q])
                except np.linalg.LinAlgError:
                    active[idx[q]] = False
        trial = p[idx] + step
        trial[:, 0] = np.maximum(trial[:, 0], 0.0)      # bounds of the reference's curve_fit
        trial[:, 3] = np.maximum(trial[:, 3], 1e-12)
        trial[:, 4] = np.maximum(trial[:, 4], 0.0)
        with np.errstate(over="ignore", invalid="ignore"):
            mt, Jt = model_jac(trial)
            rt = d[idx] - mt
            ct = (rt * rt).sum(1)
        better = np.isfinite(ct) & (ct <= cost[idx])
        bi = idx[better]
        small = np.abs(step[better]).max(1) < 1e-10 * (1.0 + np.abs(p[bi]).max(1))
        flat = (cost[bi] - ct[better]) <= 1e-14 * (cost[bi] + 1e-300)
        p[bi] = trial[better]
        J[bi], res[bi], cost[bi] = Jt[better], rt[better], ct[better]
        lam[bi] = np.maximum(lam[bi] * 0.3, 1e-12)
        done = bi[small | flat]
        converged[done] = True
        active[done] = False
        wi = idx[~better]
        lam[wi] *= 10.0
        stuck = wi[lam[wi] > 1e12]
        active[stuck] = False
    interior = (p[:, 0] > 0) & (p[:, 4] > 0) & (p[:, 3] > 1e-6)
    ok = converged & interior & np.isfinite(p).all(1)
    return p[:, 1], p[:, 2], ok


def _window_geometry(XCorr, box):
    """arg-max and fit-window cut-out of one correlation image (imageprocess.py:103-116):
    returns (y_max, x_max, FitROI or None when the window touches the crop edge)."""
    fit_X = int(box / 2)
    y_max_, x_max_ = np.unravel_index(XCorr.argmax(), XCorr.shape)
    FitROI = XCorr[y_max_ - fit_X: y_max_ + fit_X + 1, x_max_ - fit_X: x_max_ + fit_X + 1]
    dims = FitROI.shape
    if 0 in dims or dims[0] != dims[1]:
        return y_max_, x_max_, None
    return y_max_, x_max_, FitROI


def _shift_from_window(XCorr, Y, X, Y_, X_, box):
    """arg-max, fit-window cut-out and peak fit (reference imageprocess.py:103-157)."""
    y_max_, x_max_, FitROI = _window_geometry(XCorr, box)
    if FitROI is None:
        xc, yc = 0, 0
    else:
        xc, yc = _gauss_peak_fit(FitROI)
        xc += X_ + x_max_
        yc += Y_ + y_max_
        xc -= np.floor(X / 2)
        yc -= np.floor(Y / 2)
    return -yc, -xc


def get_image_shift(imageA, imageB, box: int, roi: int | None = None, display: bool = False):
    """Shift from ``imageA`` to ``imageB`` (reference imageprocess.py:53-157);
    ``(0, 0)`` if either image sums to zero or the fit window touches the crop edge."""
    imageA = np.asarray(imageA)
    imageB = np.asarray(imageB)
    if (np.sum(imageA) == 0) or (np.sum(imageB) == 0):
        return 0, 0
    win, _, (Y_, X_) = _windows(np.stack([imageA, imageB]), roi)
    Y, X = imageA.shape
    return _shift_from_window(win[0].astype(np.float64), Y, X, Y_, X_, box)


def _pair_shifts_from_windows(win, sums, pairs, Y, X, Y_, X_, callback=None):
    """(shift_y, shift_x) of every pair in ``pairs`` from its correlation window (the body of the
    reference's ``get_image_shift``, imageprocess.py:103-157, for all pairs at once): arg-max,
    5x5 window, Gaussian peak fit.  The fits of all pairs are done in one vectorised pass;
    windows it cannot settle go through scipy's ``curve_fit``."""
    geo, rois, which = [], [], []
    for flag, (i, j) in enumerate(pairs):
        if sums[i] == 0 or sums[j] == 0:
            geo.append(None)
            continue
        ym, xm, roi = _window_geometry(win[flag].astype(np.float64), 5)
        geo.append((ym, xm, roi is not None))
        if roi is not None:
            rois.append(roi)
            which.append(flag)
    if rois and rois[0].shape == (5, 5):
        bx, by, ok = _gauss_peak_fit_batch(np.stack(rois))
    else:
        bx = by = np.zeros(len(rois))
        ok = np.zeros(len(rois), bool)
    fitted = {}
    for q, flag in enumerate(which):
        fitted[flag] = (bx[q], by[q]) if ok[q] else _gauss_peak_fit(rois[q])
    sy_all = np.zeros(len(pairs))
    sx_all = np.zeros(len(pairs))
    for flag in range(len(pairs)):
        g = geo[flag]
        if g is None or not g[2]:
            sy, sx = 0, 0
        else:
            xc, yc = fitted[flag]
            xc = xc + X_ + g[1] - np.floor(X / 2)
            yc = yc + Y_ + g[0] - np.floor(Y / 2)
            sy, sx = -yc, -xc
        sy_all[flag], sx_all[flag] = sy, sx
        if callback is not None:
            callback(flag + 1)
    return sy_all, sx_all


def _rcc_from_windows(win, sums, Y, X, Y_, X_, callback):
    """Pairwise shifts from the correlation windows -> minimize_shifts (the loop body of the
    reference's rcc, imageprocess.py:191-217)."""
    n_segments = len(sums)
    shifts_x = np.zeros((n_segments, n_segments))
    shifts_y = np.zeros((n_segments, n_segments))
    n_pairs = int(n_segments * (n_segments - 1) / 2)
    bar = None
    if callback is None:
        from tqdm import tqdm

        bar = tqdm(total=n_pairs, desc="Correlating image pairs", unit="pairs")
        step = lambda k: bar.update()   # noqa: E731
    else:
        callback(0)
        step = callback
    pairs = [(i, j) for i in range(n_segments - 1) for j in range(i + 1, n_segments)]
    sy_all, sx_all = _pair_shifts_from_windows(win, sums, pairs, Y, X, Y_, X_, step)
    for flag, (i, j) in enumerate(pairs):
        shifts_y[i, j], shifts_x[i, j] = sy_all[flag], sx_all[flag]
    if bar is not None:
        bar.close()
    return lib.minimize_shifts(shifts_x, shifts_y)


def rcc(segments, max_shift: float | None = None, callback: Callable[[int], None] | None = None):
    """Redundant cross-correlation over all segment pairs (reference
    imageprocess.py:160-217); returns ``lib.minimize_shifts(shifts_x, shifts_y)`` =
    ``(shift_y, shift_x)`` per segment.  ``callback`` is called with 0..n_pairs."""
    segments = np.asarray(segments)
    n_segments = len(segments)
    if n_segments < 2:
        if callback is not None:
            callback(0)
        z = np.zeros(max(n_segments, 1))
        return lib.minimize_shifts(np.zeros((n_segments, n_segments)), np.zeros((n_segments, n_segments))) \
            if n_segments else (z[:0], z[:0])
    win, sums, (Y_, X_) = _windows(segments, max_shift)
    _, Y, X = segments.shape
    return _rcc_from_windows(win, sums, Y, X, Y_, X_, callback)


def _windows_of_locs(locs, info, bounds, min_blur_width, max_shift, pairs=None):
    """Fused device path used by postprocess.undrift: render the segment images on the GPU
    and cross-correlate them there (same kernels as segment() + rcc(), no host round trip
    of the (n_seg, Y, X) image stack).  ``pairs`` = (pair_i, pair_j) restricts the work to a
    subset of the pairs (multi-GPU sharding); returns ``(windows, sums, (Y, X, Y_, X_))``."""
    l = _lib.load()
    _declare(l)
    _lib.require_gpu()
    Y, X = info[0]["Height"], info[0]["Width"]
    n_seg = len(bounds) - 1
    seg_start, x, y, lpx, lpy = _segment_arrays(locs, info, bounds)
    Y_, X_, H, W = _crop_geometry(Y, X, max_shift)
    sums = np.zeros(n_seg, dtype=np.float64)
    if pairs is None:
        n_pairs = n_seg * (n_seg - 1) // 2
        win = np.zeros((n_pairs, H, W), dtype=np.float32)
        _lib.check(l.pb_undrift_windows(n_seg, _lib.ptr(seg_start), _lib.ptr(x), _lib.ptr(y),
                                        _lib.ptr(lpx), _lib.ptr(lpy), Y, X, float(min_blur_width),
                                        Y_, X_, H, W, _lib.ptr(win), _lib.ptr(sums), None))
    else:
        pi = np.ascontiguousarray(pairs[0], dtype=np.int32)
        pj = np.ascontiguousarray(pairs[1], dtype=np.int32)
        win = np.zeros((len(pi), H, W), dtype=np.float32)
        _lib.check(l.pb_undrift_windows_pairs(n_seg, _lib.ptr(seg_start), _lib.ptr(x), _lib.ptr(y),
                                              _lib.ptr(lpx), _lib.ptr(lpy), Y, X, float(min_blur_width),
                                              Y_, X_, H, W, len(pi), _lib.ptr(pi), _lib.ptr(pj),
                                              _lib.ptr(win), _lib.ptr(sums), None))
    return win, sums, (Y, X, Y_, X_)


def _segment_arrays(locs, info, bounds):
    """Localization columns grouped by segment (stable), the way pb_undrift_* expects them."""
    n_seg = len(bounds) - 1
    frames = locs["frame"].to_numpy()
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)   # noqa: E731
    if len(frames) < 2 or bool(np.all(frames[1:] >= frames[:-1])):
        # frames already in order (the normal case): segments are contiguous slices
        cut = np.searchsorted(frames, np.asarray(bounds), side="left")
        a, b = int(cut[0]), int(cut[-1])
        seg_start = (cut - cut[0]).astype(np.int64)
        return (seg_start,) + tuple(f32(locs[c].to_numpy()[a:b]) for c in ("x", "y", "lpx", "lpy"))
    seg_of = np.searchsorted(bounds, frames, side="right") - 1
    keep = (frames >= bounds[0]) & (frames < bounds[-1])
    idx = np.flatnonzero(keep)
    idx = idx[np.argsort(seg_of[idx], kind="stable")]
    counts = np.bincount(seg_of[idx], minlength=n_seg)[:n_seg]
    seg_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return (seg_start,) + tuple(f32(locs[c].to_numpy()[idx]) for c in ("x", "y", "lpx", "lpy"))


def _shifts_of_locs(locs, info, bounds, min_blur_width, max_shift, pairs=None, callback=None):
    """Per-pair (shift_y, shift_x) with everything up to the peak fit on the GPU
    (``pb_undrift_peaks_pairs``): segments rendered, transformed, correlated and fitted on the
    device; 256 bytes per pair come back.  Windows the device solver does not settle are re-fitted
    with scipy's bounded ``curve_fit`` like the reference; the rare square-but-not-5x5 cut-out at
    the crop corner goes through the window path."""
    l = _lib.load()
    _declare(l)
    _lib.require_gpu()
    Y, X = info[0]["Height"], info[0]["Width"]
    n_seg = len(bounds) - 1
    seg_start, x, y, lpx, lpy = _segment_arrays(locs, info, bounds)
    Y_, X_, H, W = _crop_geometry(Y, X, max_shift)
    if pairs is None:
        pi, pj = np.triu_indices(n_seg, 1)
    else:
        pi, pj = pairs
    pi = np.ascontiguousarray(pi, dtype=np.int32)
    pj = np.ascontiguousarray(pj, dtype=np.int32)
    rec = np.zeros((len(pi), 32), dtype=np.float64)
    sums = np.zeros(n_seg, dtype=np.float64)
    _lib.check(l.pb_undrift_peaks_pairs(n_seg, _lib.ptr(seg_start), _lib.ptr(x), _lib.ptr(y), _lib.ptr(lpx),
                                        _lib.ptr(lpy), Y, X, float(min_blur_width), Y_, X_, H, W, len(pi),
                                        _lib.ptr(pi), _lib.ptr(pj), _lib.ptr(rec), _lib.ptr(sums), None))
    def odd_windows(odd):
        win, _, _ = _windows_of_locs(locs, info, bounds, min_blur_width, max_shift, pairs=(pi[odd], pj[odd]))
        return win

    sy, sx = _shifts_from_records(rec, sums, pi, pj, Y, X, Y_, X_, odd_windows)
    if callback is not None:
        for flag in range(len(pi)):
            callback(flag + 1)
    return sy, sx


def _shifts_from_records(rec, sums, pi, pj, Y, X, Y_, X_, odd_windows=None):
    """(shift_y, shift_x) per pair from the device peak-fit records (32 float64 per pair: status,
    arg-max y / x, xc, yc, the 5 x 5 window) -- the tail of the reference's ``get_image_shift``
    (imageprocess.py:83-84, 113-157).  status 1: the device solver did not settle -> scipy's bounded
    ``curve_fit`` like the reference; status 2: window touches the crop edge -> (0, 0); status 3: the
    rare square-but-not-5x5 cut-out at the crop corner -> ``odd_windows(indices)`` must return those
    pairs' full windows for the host path."""
    status = rec[:, 0].astype(np.int64)
    xc, yc = rec[:, 3].copy(), rec[:, 4].copy()
    for k in np.flatnonzero(status == 1):                       # re-fit like the reference
        xc[k], yc[k] = _gauss_peak_fit(rec[k, 5:30].reshape(5, 5))
    odd = np.flatnonzero(status == 3)
    odd_shift = {}
    if len(odd):
        if odd_windows is None:
            raise RuntimeError("peak-fit status 3 needs the full correlation windows")
        win = odd_windows(odd)
        for q, k in enumerate(odd):
            odd_shift[k] = _shift_from_window(np.asarray(win[q], dtype=np.float64), Y, X, Y_, X_, 5)
    zero = (sums[pi] == 0) | (sums[pj] == 0) | (status == 2)
    sx = -(xc + X_ + rec[:, 2] - np.floor(X / 2))
    sy = -(yc + Y_ + rec[:, 1] - np.floor(Y / 2))
    sx[zero] = 0
    sy[zero] = 0
    for k, (a, b) in odd_shift.items():
        if sums[pi[k]] != 0 and sums[pj[k]] != 0:
            sy[k], sx[k] = a, b
    return sy, sx


def _rcc_of_locs(locs, info, bounds, min_blur_width, max_shift, callback):
    """Segment shifts for postprocess.undrift: render, cross-correlate and peak-fit on the GPU,
    then ``minimize_shifts`` (reference imageprocess.py:160-217)."""
    n_seg = len(bounds) - 1
    bar = None
    if callback is None:
        from tqdm import tqdm

        bar = tqdm(total=n_seg * (n_seg - 1) // 2, desc="Correlating image pairs", unit="pairs")
        step = lambda k: bar.update()   # noqa: E731
    else:
        callback(0)
        step = callback
    sy, sx = _shifts_of_locs(locs, info, bounds, min_blur_width, max_shift, callback=step)
    if bar is not None:
        bar.close()
    shifts_x = np.zeros((n_seg, n_seg))
    shifts_y = np.zeros((n_seg, n_seg))
    pi, pj = np.triu_indices(n_seg, 1)
    shifts_y[pi, pj] = sy
    shifts_x[pi, pj] = sx
    return lib.minimize_shifts(shifts_x, shifts_y)


def _rcc_of_locs_windows(locs, info, bounds, min_blur_width, max_shift, callback):
    """Same through the window path (host peak fits); kept for A/B tests."""
    win, sums, (Y, X, Y_, X_) = _windows_of_locs(locs, info, bounds, min_blur_width, max_shift)
    return _rcc_from_windows(win, sums, Y, X, Y_, X_, callback)
