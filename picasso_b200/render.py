"""Super-resolution rendering on B200.

Drop-in for ``picasso.render.render`` (reference picasso/render.py:37-174) for the
unrotated ``None`` (histogram), ``"gaussian"`` and ``"gaussian_iso"`` blur methods --
the ones the localization / undrift hot path uses.  ``"smooth"``, ``"convolve"`` and
rotated rendering (``ang``) are outside the B200 hot path
and raise ``NotImplementedError``.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib, lib

_MODES = {None: 0, "gaussian": 1, "gaussian_iso": 2}


def _declare(l):
    if getattr(l, "_render_declared", False):
        return
    vp, i32, f64, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    l.pb_render.argtypes = [sz, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, i32, vp, i32, i32,
                            C.POINTER(C.c_longlong)]
    l.pb_render.restype = i32
    l.pb_render_workspace_bytes.argtypes = [sz, i32, i32]
    l.pb_render_workspace_bytes.restype = sz
    l.pb_render_dev.argtypes = [sz, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, i32, vp, i32,
                                i32, vp, vp, sz, vp]
    l.pb_render_dev.restype = i32
    l._render_declared = True


def render(locs, info, oversampling: float = 1.0, viewport=None, blur_method=None,
           min_blur_width: float = 0.0, ang=None, disp_px_size: float | None = None):
    """Render localizations into an image; returns ``(n, image)``.

    Same contract as the reference: ``Pixelsize`` must be in ``info`` (KeyError);
    ``oversampling`` is deprecated in favour of ``disp_px_size``; the default viewport
    is the full field of view from ``info[0]``; unknown ``blur_method`` raises
    ``Exception("blur_method not understood.")``; ``image`` is float32 with shape
    ``(ceil(os * dy), ceil(os * dx))`` and ``n`` counts localizations strictly inside
    the viewport.
    """
    pixelsize = lib.get_from_metadata(info, "Pixelsize", raise_error=True)
    if disp_px_size is None:
        warnings.warn(
            "Deprecation warning: the 'oversampling' parameter is deprecated and will be "
            "removed in v0.11.0. Use 'disp_px_size' instead.", DeprecationWarning, stacklevel=2)
        disp_px_size = pixelsize / oversampling
    oversampling = pixelsize / disp_px_size
    if viewport is None:
        try:
            viewport = [(0, 0), (info[0]["Height"], info[0]["Width"])]
        except TypeError:
            raise ValueError("Need info if no viewport is provided.")
    (y_min, x_min), (y_max, x_max) = viewport
    if blur_method not in _MODES:
        if blur_method in ("smooth", "convolve"):
            raise NotImplementedError(
                f"blur_method={blur_method!r} is outside the B200 hot path; use the reference")
        raise Exception("blur_method not understood.")
    if ang is not None:
        raise NotImplementedError("rotated rendering (ang=...) is outside the B200 hot path")
    mode = _MODES[blur_method]
    l = _lib.load()
    _declare(l)
    _lib.require_gpu()
    n_pixel_y = int(np.ceil(oversampling * (y_max - y_min)))
    n_pixel_x = int(np.ceil(oversampling * (x_max - x_min)))
    x = np.ascontiguousarray(locs["x"], dtype=np.float32)
    y = np.ascontiguousarray(locs["y"], dtype=np.float32)
    lpx = lpy = None
    if mode:
        lpx = np.ascontiguousarray(locs["lpx"], dtype=np.float32)
        lpy = np.ascontiguousarray(locs["lpy"], dtype=np.float32)
    image = np.zeros((max(n_pixel_y, 0), max(n_pixel_x, 0)), dtype=np.float32)
    n = C.c_longlong(0)
    _lib.check(l.pb_render(len(x), _lib.ptr(x), _lib.ptr(y),
                           _lib.ptr(lpx) if mode else None, _lib.ptr(lpy) if mode else None,
                           float(oversampling), float(y_min), float(x_min), float(y_max),
                           float(x_max), float(min_blur_width), mode, _lib.ptr(image),
                           n_pixel_y, n_pixel_x, C.byref(n)))
    return int(n.value), image
