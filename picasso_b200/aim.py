"""Adaptive Intersection Maximization (AIM) drift correction on B200.

Drop-in for ``picasso.aim`` (reference picasso/aim.py; Ma et al., Science Advances 2024):
``aim`` :776, ``intersection_max`` :517, ``intersection_max_z`` :662, ``get_fft_peak`` :434 and
``get_fft_peak_z`` :480 keep their signatures and results.  The intersection counting -- the
reference sorts the concatenated coordinate arrays once per shift and segment -- runs on the
GPU against a resident hash table (csrc/aim.cu through ``pb_aim_*``) and is bit-exact; the
sub-pixel peak (phase of the first Fourier coefficients of the small count array), the cubic
spline and the subtraction are the reference's numpy / scipy calls on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Literal

import numpy as np
import pandas as pd
from scipy.interpolate import InterpolatedUnivariateSpline

from . import __version__, _lib, lib


def _declare(l):
    if getattr(l, "_aim_declared", False):
        return
    vp, i32, f64, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    l.pb_aim_create.argtypes = [C.POINTER(vp)]
    l.pb_aim_destroy.argtypes = [vp]
    l.pb_aim_set_targets.argtypes = [vp, sz, vp, i32, vp, i32, vp, i32]
    l.pb_aim_set_reference.argtypes = [vp, sz, vp, i32, vp, i32, vp, i32, f64, f64, f64]
    l.pb_aim_count.argtypes = [vp, sz, sz, f64, f64, f64, i32, vp, vp]
    for f in (l.pb_aim_create, l.pb_aim_destroy, l.pb_aim_set_targets, l.pb_aim_set_reference,
              l.pb_aim_count):
        f.restype = i32
    l._aim_declared = True


def _coord(a):
    """Contiguous float32 / float64 array (other dtypes promote to float64) and its flag."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        return np.ascontiguousarray(a), 0
    return np.ascontiguousarray(a, dtype=np.float64), 1


class _Counter:
    """One AIM round on the GPU: targets in frame order + reference table."""

    def __init__(self):
        self.l = _lib.load()
        _declare(self.l)
        _lib.require_gpu()
        self.h = C.c_void_p()
        _lib.check(self.l.pb_aim_create(C.byref(self.h)))

    def close(self):
        if self.h:
            self.l.pb_aim_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_targets(self, x, y, z=None):
        x, fx = _coord(x)
        y, fy = _coord(y)
        z, fz = _coord(z) if z is not None else (None, 0)
        _lib.check(self.l.pb_aim_set_targets(self.h, len(x), _lib.ptr(x), fx, _lib.ptr(y), fy,
                                             _lib.ptr(z) if z is not None else None, fz))

    def set_reference(self, rx, ry, rz, intersect_d, width_units, height_units=0.0):
        rx, fx = _coord(rx)
        ry, fy = _coord(ry)
        rz, fz = _coord(rz) if rz is not None else (None, 0)
        _lib.check(self.l.pb_aim_set_reference(self.h, len(rx), _lib.ptr(rx), fx, _lib.ptr(ry), fy,
                                               _lib.ptr(rz) if rz is not None else None, fz,
                                               float(intersect_d), float(width_units), float(height_units)))

    def count(self, first, count, rel_x, rel_y, rel_z, shifts):
        shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        roi = np.zeros(len(shifts), np.int32)
        _lib.check(self.l.pb_aim_count(self.h, int(first), int(count), float(rel_x), float(rel_y),
                                       float(rel_z), len(shifts), _lib.ptr(shifts), _lib.ptr(roi)))
        return roi


def _get_fft_peak(roi_cc, roi_size):
    """Sub-pixel peak of the 2-D count array (reference aim.py:444-477)."""
    fft_values = np.fft.fft2(roi_cc.T)
    ang_x = np.angle(fft_values[0, 1])
    ang_x = ang_x - 2 * np.pi * (ang_x > 0)
    px = np.abs(ang_x) / (2 * np.pi / roi_cc.shape[0]) - (roi_cc.shape[0] - 1) / 2
    px *= roi_size / roi_cc.shape[0]
    ang_y = np.angle(fft_values[1, 0])
    ang_y = ang_y - 2 * np.pi * (ang_y > 0)
    py = np.abs(ang_y) / (2 * np.pi / roi_cc.shape[1]) - (roi_cc.shape[1] - 1) / 2
    py *= roi_size / roi_cc.shape[1]
    return px, py


def _get_fft_peak_z(roi_cc, roi_size):
    """Sub-pixel peak of the 1-D z count array (reference aim.py:490-514)."""
    fft_values = np.fft.fft(roi_cc)
    ang_z = np.angle(fft_values[1])
    ang_z = ang_z - 2 * np.pi * (ang_z > 0)
    pz = np.abs(ang_z) / (2 * np.pi / roi_cc.size) - (roi_cc.size - 1) / 2
    pz *= roi_size / roi_cc.size
    return pz


get_fft_peak = _get_fft_peak
get_fft_peak_z = _get_fft_peak_z


def _segments_in_frame_order(frame, seg_bounds):
    """Stable frame order and, per segment s, the range of sorted positions with
    seg_bounds[s] < frame <= seg_bounds[s + 1] (the reference's boolean masks)."""
    f = np.asarray(frame)
    if len(f) < 2 or bool(np.all(f[1:] >= f[:-1])):
        order, fs = slice(None), f            # already in frame order (the normal case): no gather
    else:
        order = np.argsort(f, kind="stable")
        fs = f[order]
    start = np.searchsorted(fs, seg_bounds[:-1], side="right")
    end = np.searchsorted(fs, seg_bounds[1:], side="right")
    return order, start, end


def _iterator(progress, start, n):
    if progress is not None and hasattr(progress, "get_iterator"):
        return progress.get_iterator(start, n)
    return range(start, n)


def _spline(seg_bounds, d):
    t = (seg_bounds[1:] + seg_bounds[:-1]) / 2
    return InterpolatedUnivariateSpline(t, d, k=3)(np.arange(seg_bounds[-1]) + 1)


def intersection_max(x, y, ref_x, ref_y, frame, seg_bounds, intersect_d, roi_r, width, aim_round=1,
                     progress=None, _record=None):
    """Undrift x, y against the reference by intersection maximization (reference
    ``intersection_max``, aim.py:517-659): returns ``(x_pdc, y_pdc, drift_x, drift_y)``."""
    assert aim_round in [1, 2], "aim_round must be 1 or 2."
    n_segments = len(seg_bounds) - 1
    rel_drift_x = 0
    rel_drift_y = 0
    drift_x = np.zeros(n_segments)
    drift_y = np.zeros(n_segments)
    roi_units = int(np.ceil(roi_r / intersect_d))
    steps = np.arange(-roi_units, roi_units + 1, 1)
    box = len(steps)
    shifts_xy = np.zeros((box, box), dtype=np.int32)
    width_units = width / intersect_d
    for i, shift_x in enumerate(steps):
        for j, shift_y in enumerate(steps):
            shifts_xy[i, j] = shift_x + shift_y * width_units
    shifts_xy = shifts_xy.reshape(box ** 2)
    order, seg_start, seg_end = _segments_in_frame_order(frame, seg_bounds)
    with _Counter() as gpu:
        gpu.set_targets(np.asarray(x)[order], np.asarray(y)[order])
        gpu.set_reference(ref_x, ref_y, None, intersect_d, width_units)
        for s in _iterator(progress, 1 if aim_round == 1 else 0, n_segments):
            n1 = int(seg_end[s] - seg_start[s])
            if n1 == 0:
                drift_x[s] = drift_x[s - 1]
                drift_y[s] = drift_y[s - 1]
                continue
            roi_cc = gpu.count(seg_start[s], n1, rel_drift_x, rel_drift_y, 0.0, shifts_xy).reshape(box, box)
            if _record is not None:
                _record.append(roi_cc)
            px, py = _get_fft_peak(roi_cc, 2 * roi_r)
            rel_drift_x += px
            rel_drift_y += py
            drift_x[s] = -rel_drift_x
            drift_y[s] = -rel_drift_y
            if progress is not None and hasattr(progress, "set_value"):
                progress.set_value(s)
    drift_x = _spline(seg_bounds, drift_x)
    drift_y = _spline(seg_bounds, drift_y)
    x_pdc = x - drift_x[frame - 1]
    y_pdc = y - drift_y[frame - 1]
    return x_pdc, y_pdc, drift_x, drift_y


def intersection_max_z(x, y, z, ref_x, ref_y, ref_z, frame, seg_bounds, intersect_d, roi_r, width,
                       height, pixelsize, aim_round=1, progress=None, _record=None):
    """Undrift z (nm) for already undrifted x, y (reference ``intersection_max_z``,
    aim.py:662-773): returns ``(z_pdc, drift_z)`` in nm."""
    assert aim_round in [1, 2], "aim_round must be 1 or 2."
    z = z.copy() / pixelsize
    ref_z = ref_z.copy() / pixelsize
    n_segments = len(seg_bounds) - 1
    rel_drift_z = 0
    drift_z = np.zeros(n_segments)
    roi_units = int(np.ceil(roi_r / intersect_d))
    steps = np.arange(-roi_units, roi_units + 1, 1)
    width_units = width / intersect_d
    height_units = height / intersect_d
    shifts_z = steps.astype(np.int32) * width_units * height_units
    order, seg_start, seg_end = _segments_in_frame_order(frame, seg_bounds)
    with _Counter() as gpu:
        gpu.set_targets(np.asarray(x)[order], np.asarray(y)[order], np.asarray(z)[order])
        gpu.set_reference(ref_x, ref_y, ref_z, intersect_d, width_units, height_units)
        for s in _iterator(progress, 1 if aim_round == 1 else 0, n_segments):
            n1 = int(seg_end[s] - seg_start[s])
            if n1 == 0:
                drift_z[s] = drift_z[s - 1]
                continue
            roi_cc = gpu.count(seg_start[s], n1, 0.0, 0.0, rel_drift_z, shifts_z)
            if _record is not None:
                _record.append(roi_cc)
            pz = _get_fft_peak_z(roi_cc, 2 * roi_r)
            rel_drift_z += pz
            drift_z[s] = -rel_drift_z
            if progress is not None and hasattr(progress, "set_value"):
                progress.set_value(s)
    drift_z = _spline(seg_bounds, drift_z)
    z_pdc = z - drift_z[frame - 1]
    z_pdc *= pixelsize
    drift_z *= pixelsize
    return z_pdc, drift_z


def aim(locs: pd.DataFrame, info: list[dict], segmentation: int = 100, intersect_d: float = 20 / 130,
        roi_r: float = 60 / 130, progress=None):
    """Apply AIM undrifting (reference ``aim``, aim.py:776-950): returns
    ``(locs, new_info, drift)`` with ``drift`` a float32 DataFrame (x, y[, z]) per frame."""
    assert progress is None or progress == "console" or hasattr(progress, "get_iterator"), (
        "progress must be None, 'console', or a ProgressDialog instance.")
    if progress == "console":
        progress = None
    locs = locs.copy()
    width = lib.get_from_metadata(info, "Width", raise_error=True)
    height = lib.get_from_metadata(info, "Height", raise_error=True)
    pixelsize = lib.get_from_metadata(info, "Pixelsize", raise_error=True)
    n_frames = lib.get_from_metadata(info, "Frames", raise_error=True)
    frame = locs["frame"] + 1 - locs["frame"].min()
    seg_bounds = np.concatenate((np.arange(0, n_frames, segmentation), [n_frames]))
    ref_x = locs["x"][frame <= segmentation]
    ref_y = locs["y"][frame <= segmentation]
    x_pdc, y_pdc, drift_x1, drift_y1 = intersection_max(
        locs["x"], locs["y"], ref_x, ref_y, frame, seg_bounds, intersect_d, roi_r, width, aim_round=1,
        progress=progress)
    if progress is not None and hasattr(progress, "zero_progress"):
        progress.zero_progress(description="Undrifting by AIM (2/2)")
    x_pdc, y_pdc, drift_x2, drift_y2 = intersection_max(
        x_pdc, y_pdc, x_pdc, y_pdc, frame, seg_bounds, intersect_d, roi_r, width, aim_round=2,
        progress=progress)
    drift_x = drift_x1 + drift_x2
    drift_y = drift_y1 + drift_y2
    shift_x = np.mean(drift_x)
    shift_y = np.mean(drift_y)
    drift_x -= shift_x
    drift_y -= shift_y
    x_pdc += shift_x
    y_pdc += shift_y
    if "z" in locs.columns:
        ref_x = x_pdc[frame <= segmentation]
        ref_y = y_pdc[frame <= segmentation]
        ref_z = locs["z"][frame <= segmentation]
        z_pdc, drift_z1 = intersection_max_z(
            x_pdc, y_pdc, locs["z"], ref_x, ref_y, ref_z, frame, seg_bounds, intersect_d, roi_r, width,
            height, pixelsize, aim_round=1, progress=progress)
        z_pdc, drift_z2 = intersection_max_z(
            x_pdc, y_pdc, z_pdc, x_pdc, y_pdc, z_pdc, frame, seg_bounds, intersect_d, roi_r, width,
            height, pixelsize, aim_round=2, progress=progress)
        drift_z = drift_z1 + drift_z2
        shift_z = np.mean(drift_z)
        drift_z -= shift_z
        z_pdc += shift_z
        drift = pd.DataFrame({"x": drift_x, "y": drift_y, "z": drift_z}, dtype="float32")
    else:
        drift = pd.DataFrame({"x": drift_x, "y": drift_y}, dtype="float32")
    locs["x"] = x_pdc
    locs["y"] = y_pdc
    if "z" in locs.columns:
        locs["z"] = z_pdc
    new_info = {
        "Generated by": f"Picasso v{__version__} AIM (picasso_b200)",
        "Intersect distance (nm)": intersect_d * pixelsize,
        "Segmentation": segmentation,
        "Search regions radius (nm)": roi_r * pixelsize,
    }
    if progress is not None and hasattr(progress, "close"):
        progress.close()
    return locs, info + [new_info], drift
