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
    l.pb_copy_h2d.argtypes = [vp, vp, sz, vp]
    l.pb_copy_h2d.restype = i32
    l.pb_copy_d2h.argtypes = [vp, vp, sz, vp]
    l.pb_copy_d2h.restype = i32
    l._render_declared = True


def _setup(info, oversampling, viewport, blur_method, ang, disp_px_size):
    """Argument handling shared by ``render`` and the device-resident variant: returns
    ``(oversampling, (y_min, x_min, y_max, x_max), mode, n_pixel_y, n_pixel_x)``."""
    pixelsize = lib.get_from_metadata(info, "Pixelsize", raise_error=True)
    if disp_px_size is None:
        warnings.warn(
            "Deprecation warning: the 'oversampling' parameter is deprecated and will be "
            "removed in v0.11.0. Use 'disp_px_size' instead.", DeprecationWarning, stacklevel=3)
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
    n_pixel_y = int(np.ceil(oversampling * (y_max - y_min)))
    n_pixel_x = int(np.ceil(oversampling * (x_max - x_min)))
    return oversampling, (y_min, x_min, y_max, x_max), _MODES[blur_method], n_pixel_y, n_pixel_x


def render_to_device(torch, locs, info, oversampling: float = 1.0, viewport=None, blur_method=None,
                     min_blur_width: float = 0.0, ang=None, disp_px_size: float | None = None,
                     device="cuda"):
    """``render`` that leaves the image on the GPU: returns ``(count int64 tensor[1], image
    float32 tensor)`` on ``device`` -- the multi-GPU path all-reduces them before one download."""
    oversampling, (y_min, x_min, y_max, x_max), mode, ny, nx = _setup(
        info, oversampling, viewport, blur_method, ang, disp_px_size)
    l = _lib.load()
    _declare(l)
    _lib.require_gpu()
    st = torch.cuda.current_stream(device).cuda_stream
    cols = ["x", "y"] + (["lpx", "lpy"] if mode else [])
    n = len(locs["x"])
    dev = {}
    for c in cols:
        h = np.ascontiguousarray(locs[c], dtype=np.float32)
        t = torch.empty(n, dtype=torch.float32, device=device)
        _lib.check(l.pb_copy_h2d(t.data_ptr(), _lib.ptr(h), n * 4, st))
        dev[c] = t
    image = torch.empty((max(ny, 0), max(nx, 0)), dtype=torch.float32, device=device)
    count = torch.zeros(1, dtype=torch.int64, device=device)
    wsb = l.pb_render_workspace_bytes(n, ny, nx) if mode else 0
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=device)
    if image.numel():
        _lib.check(l.pb_render_dev(n, dev["x"].data_ptr(), dev["y"].data_ptr(),
                                   dev["lpx"].data_ptr() if mode else None,
                                   dev["lpy"].data_ptr() if mode else None, float(oversampling),
                                   float(y_min), float(x_min), float(y_max), float(x_max),
                                   float(min_blur_width), mode, image.data_ptr(), ny, nx,
                                   count.data_ptr(), ws.data_ptr() if mode else None, wsb, st))
    return count, image


def image_to_host(torch, image):
    """Download a device image through the threaded pinned staging (pageable numpy result)."""
    l = _lib.load()
    _declare(l)
    out = _lib.pinned_empty(tuple(image.shape), np.float32)
    if out.size:
        st = torch.cuda.current_stream(image.device).cuda_stream
        _lib.check(l.pb_copy_d2h(_lib.ptr(out), image.data_ptr(), out.nbytes, st))
    return out


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
    # pb_render writes every pixel; page-locked backing makes the 419 MB download of a 10240^2
    # image one direct DMA
    image = _lib.pinned_empty((max(n_pixel_y, 0), max(n_pixel_x, 0)), np.float32)
    if image.size == 0 or len(x) == 0:
        image[...] = 0
    n = C.c_longlong(0)
    _lib.check(l.pb_render(len(x), _lib.ptr(x), _lib.ptr(y),
                           _lib.ptr(lpx) if mode else None, _lib.ptr(lpy) if mode else None,
                           float(oversampling), float(y_min), float(x_min), float(y_max),
                           float(x_max), float(min_blur_width), mode, _lib.ptr(image),
                           n_pixel_y, n_pixel_x, C.byref(n)))
    return int(n.value), image
