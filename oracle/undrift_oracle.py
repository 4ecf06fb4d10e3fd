"""oracle/undrift_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's RCC drift correction:
  picasso/imageprocess.py @ 96e0da51: xcorr :27-50, get_image_shift :53-157, rcc :160-217
  picasso/lib.py: minimize_shifts :2034-2078
  picasso/postprocess.py: n_segments :2824-2843, segment :2846-2900, undrift :2903-2961,
                          _apply_drift :3159-3168
Third-party arithmetic is the same the reference calls: numpy.fft (pocketfft, float64),
scipy.optimize.curve_fit (bounded -> least_squares 'trf'), scipy InterpolatedUnivariateSpline,
numpy.linalg.pinv.  Segment images come from the render oracle (render_oracle.c).
Pinned by tests/test_oracle_golden_undrift.py against outputs of the real reference.
"""
from __future__ import annotations

import numpy as np


def xcorr(a, b):
    """imageprocess.py:27-50"""
    fa = np.fft.fft2(a)
    cfb = np.conj(np.fft.fft2(b))
    return np.fft.fftshift(np.real(np.fft.ifft2(fa * cfb))) / np.sqrt(a.size)


def get_image_shift(a, b, box, roi=None):
    """imageprocess.py:53-157"""
    from scipy.optimize import curve_fit

    if np.sum(a) == 0 or np.sum(b) == 0:
        return 0, 0
    xc_img = xcorr(a, b)
    Y, X = a.shape
    if roi is not None:
        Y_ = int((Y - roi) / 2)
        X_ = int((X - roi) / 2)
        if Y_ > 0:
            xc_img = xc_img[Y_:-Y_, :]
        else:
            Y_ = 0
        if X_ > 0:
            xc_img = xc_img[:, X_:-X_]
        else:
            X_ = 0
    else:
        Y_ = X_ = 0
    k = int(box / 2)
    y, x = np.mgrid[-k:k + 1, -k:k + 1]
    ym, xm = np.unravel_index(xc_img.argmax(), xc_img.shape)
    fit = xc_img[ym - k: ym + k + 1, xm - k: xm + k + 1]
    if 0 in fit.shape or fit.shape[0] != fit.shape[1]:
        xc, yc = 0, 0
    else:
        def g2(coords, a_, xc_, yc_, s_, b_):
            xx, yy = coords
            return (a_ * np.exp(-0.5 * ((xx - xc_) ** 2 + (yy - yc_) ** 2) / s_ ** 2) + b_).flatten()

        p0 = [fit.max(), 0, 0, 1, fit.min()]
        bounds = ([0, -np.inf, -np.inf, 0, 0], [np.inf, np.inf, np.inf, np.inf, np.inf])
        popt, _ = curve_fit(g2, (x, y), fit.flatten(), p0=p0, bounds=bounds)
        xc = popt[1] + X_ + xm - np.floor(X / 2)
        yc = popt[2] + Y_ + ym - np.floor(Y / 2)
    return -yc, -xc


def minimize_shifts(shifts_x, shifts_y):
    """lib.py:2034-2078"""
    n = shifts_x.shape[0]
    n_pairs = int(n * (n - 1) / 2)
    rij = np.zeros((n_pairs, 2))
    A = np.zeros((n_pairs, n - 1))
    flag = 0
    for i in range(n - 1):
        for j in range(i + 1, n):
            rij[flag, 0] = shifts_y[i, j]
            rij[flag, 1] = shifts_x[i, j]
            A[flag, i:j] = 1
            flag += 1
    Dj = np.dot(np.linalg.pinv(A), rij)
    return np.insert(np.cumsum(Dj[:, 0]), 0, 0), np.insert(np.cumsum(Dj[:, 1]), 0, 0)


def pair_shifts(segments, max_shift=None):
    n = len(segments)
    sx = np.zeros((n, n))
    sy = np.zeros((n, n))
    for i in range(n - 1):
        for j in range(i + 1, n):
            sy[i, j], sx[i, j] = get_image_shift(segments[i], segments[j], 5, max_shift)
    return sy, sx


def rcc(segments, max_shift=None):
    """imageprocess.py:160-217"""
    sy, sx = pair_shifts(segments, max_shift)
    return minimize_shifts(sx, sy)


def segment(locs, info, segmentation, render_fn, kwargs):
    """postprocess.py:2846-2900 (render_fn = the render oracle)"""
    Y, X, n_frames = info[0]["Height"], info[0]["Width"], info[0]["Frames"]
    n_seg = int(np.round(n_frames / segmentation))
    bounds = np.linspace(0, n_frames - 1, n_seg + 1, dtype=np.uint32)
    segments = np.zeros((n_seg, Y, X))
    fr = np.asarray(locs["frame"])
    for i in range(n_seg):
        sel = (fr >= bounds[i]) & (fr < bounds[i + 1])
        sub = {k: np.asarray(locs[k])[sel] for k in ("x", "y", "lpx", "lpy")}
        _, segments[i] = render_fn(sub, info, **kwargs)
    return bounds, segments


def undrift(locs, info, segmentation, render_fn):
    """postprocess.py:2903-2961; returns (drift (Frames, 2) [x, y], x_new, y_new)."""
    from scipy import interpolate

    bounds, segments = segment(locs, info, segmentation, render_fn,
                               {"blur_method": "gaussian", "min_blur_width": 1})
    shift_y, shift_x = rcc(segments, 32)
    t = (bounds[1:] + bounds[:-1]) / 2
    px = interpolate.InterpolatedUnivariateSpline(t, shift_x, k=3)
    py = interpolate.InterpolatedUnivariateSpline(t, shift_y, k=3)
    ti = np.arange(info[0]["Frames"])
    dx, dy = px(ti), py(ti)
    fr = np.asarray(locs["frame"])
    x_new = np.asarray(locs["x"]) - dx[fr]
    y_new = np.asarray(locs["y"]) - dy[fr]
    return np.stack([dx, dy], 1), x_new, y_new
