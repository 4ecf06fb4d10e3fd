"""oracle/aim_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU restatement, never on the product path).

Adaptive Intersection Maximization drift correction, restating picasso.aim (reference
picasso/aim.py @ 96e0da51): `_count_intersections` :89-126, `_run_intersections` :148-191,
`_point_intersect_2d` :297-344, `_point_intersect_3d` :377-431, `_get_fft_peak` :444-477,
`_get_fft_peak_z` :490-514, `intersection_max` :517-659, `intersection_max_z` :662-773,
`aim` :776-950.  Arithmetic follows the reference's dtypes: pandas float32 columns stay float32
through `+= drift`, `/ intersect_d`, `np.round` and the 1-D index `x + y * width_units`
(numpy/pandas treat Python and numpy float scalars as weak), float64 after the first round.
Pinned by tests/golden/aim.npz (drift, undrifted coordinates and every per-segment
intersection-count array of the real reference).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.interpolate import InterpolatedUnivariateSpline


def count_grid(l0_coords, l0_counts, l1_coords, l1_counts, shifts):
    """Number of intersections for every shift: sum over coordinates common to the reference
    and the shifted target of min(count0, count1)  (aim.py:89-126, 148-191)."""
    out = np.zeros(len(shifts), dtype=np.int32)
    for k, sh in enumerate(shifts):
        c = l1_coords + sh
        pos = np.searchsorted(l0_coords, c)
        pos[pos >= len(l0_coords)] = 0
        hit = (l0_coords[pos] == c) if len(l0_coords) else np.zeros(len(c), bool)
        out[k] = np.sum(np.minimum(l0_counts[pos[hit]], l1_counts[hit]))
    return out


def fft_peak_2d(roi_cc, roi_size):
    """Sub-pixel peak from the phase of the first Fourier coefficients (aim.py:444-477)."""
    f = np.fft.fft2(roi_cc.T)
    n0, n1 = roi_cc.shape
    ax = np.angle(f[0, 1]); ax = ax - 2 * np.pi * (ax > 0)
    px = (np.abs(ax) / (2 * np.pi / n0) - (n0 - 1) / 2) * (roi_size / n0)
    ay = np.angle(f[1, 0]); ay = ay - 2 * np.pi * (ay > 0)
    py = (np.abs(ay) / (2 * np.pi / n1) - (n1 - 1) / 2) * (roi_size / n1)
    return px, py


def fft_peak_1d(roi_cc, roi_size):
    f = np.fft.fft(roi_cc)
    a = np.angle(f[1]); a = a - 2 * np.pi * (a > 0)
    return (np.abs(a) / (2 * np.pi / roi_cc.size) - (roi_cc.size - 1) / 2) * (roi_size / roi_cc.size)


def _spline_all_frames(seg_bounds, d):
    t = (seg_bounds[1:] + seg_bounds[:-1]) / 2
    return InterpolatedUnivariateSpline(t, d, k=3)(np.arange(seg_bounds[-1]) + 1)


def intersection_max(x, y, ref_x, ref_y, frame, seg_bounds, intersect_d, roi_r, width, aim_round=1,
                     record=None):
    n_seg = len(seg_bounds) - 1
    rel_x = rel_y = 0
    drift_x = np.zeros(n_seg); drift_y = np.zeros(n_seg)
    ru = int(np.ceil(roi_r / intersect_d))
    steps = np.arange(-ru, ru + 1, 1)
    box = len(steps)
    width_units = width / intersect_d
    shifts = np.zeros((box, box), dtype=np.int32)
    for i, sx in enumerate(steps):
        for j, sy in enumerate(steps):
            shifts[i, j] = sx + sy * width_units
    shifts = shifts.reshape(box ** 2)
    l0 = np.int32(np.round(ref_x / intersect_d) + np.round(ref_y / intersect_d) * width_units)
    l0_coords, l0_counts = np.unique(l0, return_counts=True)
    for s in range(1 if aim_round == 1 else 0, n_seg):
        sel = (frame > seg_bounds[s]) & (frame <= seg_bounds[s + 1])
        x1 = x[sel]; y1 = y[sel]
        if len(x1) == 0:
            drift_x[s] = drift_x[s - 1]; drift_y[s] = drift_y[s - 1]
            continue
        x1 += rel_x
        y1 += rel_y
        l1 = np.int32(np.round(x1 / intersect_d) + np.round(y1 / intersect_d) * width_units)
        l1_coords, l1_counts = np.unique(l1, return_counts=True)
        roi_cc = count_grid(l0_coords, l0_counts, l1_coords, l1_counts, shifts).reshape(box, box)
        if record is not None:
            record.append(roi_cc)
        px, py = fft_peak_2d(roi_cc, 2 * roi_r)
        rel_x += px; rel_y += py
        drift_x[s] = -rel_x; drift_y[s] = -rel_y
    drift_x = _spline_all_frames(seg_bounds, drift_x)
    drift_y = _spline_all_frames(seg_bounds, drift_y)
    return x - drift_x[frame - 1], y - drift_y[frame - 1], drift_x, drift_y


def intersection_max_z(x, y, z, ref_x, ref_y, ref_z, frame, seg_bounds, intersect_d, roi_r, width, height,
                       pixelsize, aim_round=1, record=None):
    z = z.copy() / pixelsize
    ref_z = ref_z.copy() / pixelsize
    n_seg = len(seg_bounds) - 1
    rel_z = 0
    drift_z = np.zeros(n_seg)
    ru = int(np.ceil(roi_r / intersect_d))
    steps = np.arange(-ru, ru + 1, 1)
    width_units = width / intersect_d
    height_units = height / intersect_d
    shifts_z = steps.astype(np.int32) * width_units * height_units
    l0 = np.int32(np.round(ref_x / intersect_d) + np.round(ref_y / intersect_d) * width_units
                  + np.round(ref_z / intersect_d) * width_units * height_units)
    l0_coords, l0_counts = np.unique(l0, return_counts=True)
    for s in range(1 if aim_round == 1 else 0, n_seg):
        sel = (frame > seg_bounds[s]) & (frame <= seg_bounds[s + 1])
        x1 = x[sel]; y1 = y[sel]; z1 = z[sel]
        if len(x1) == 0:
            drift_z[s] = drift_z[s - 1]
            continue
        z1 += rel_z
        l1 = np.int32(np.round(x1 / intersect_d) + np.round(y1 / intersect_d) * width_units
                      + np.round(z1 / intersect_d) * width_units * height_units)
        l1_coords, l1_counts = np.unique(l1, return_counts=True)
        roi_cc = count_grid(l0_coords, l0_counts, l1_coords, l1_counts, shifts_z)
        if record is not None:
            record.append(roi_cc)
        rel_z += fft_peak_1d(roi_cc, 2 * roi_r)
        drift_z[s] = -rel_z
    drift_z = _spline_all_frames(seg_bounds, drift_z)
    z_pdc = z - drift_z[frame - 1]
    z_pdc *= pixelsize
    drift_z *= pixelsize
    return z_pdc, drift_z


def aim(locs, info, segmentation=100, intersect_d=20 / 130, roi_r=60 / 130, record=None):
    """picasso.aim.aim (aim.py:776-950): returns (locs, drift DataFrame float32)."""
    locs = locs.copy()
    width, height = info[0]["Width"], info[0]["Height"]
    pixelsize, n_frames = info[0]["Pixelsize"], info[0]["Frames"]
    frame = locs["frame"] + 1 - locs["frame"].min()
    seg_bounds = np.concatenate((np.arange(0, n_frames, segmentation), [n_frames]))
    rec2 = [] if record is not None else None
    rec3 = [] if record is not None else None
    ref_x = locs["x"][frame <= segmentation]
    ref_y = locs["y"][frame <= segmentation]
    x_pdc, y_pdc, dx1, dy1 = intersection_max(locs["x"], locs["y"], ref_x, ref_y, frame, seg_bounds,
                                              intersect_d, roi_r, width, 1, rec2)
    x_pdc, y_pdc, dx2, dy2 = intersection_max(x_pdc, y_pdc, x_pdc, y_pdc, frame, seg_bounds, intersect_d,
                                              roi_r, width, 2, rec2)
    drift_x = dx1 + dx2; drift_y = dy1 + dy2
    sx, sy = np.mean(drift_x), np.mean(drift_y)
    drift_x -= sx; drift_y -= sy
    x_pdc += sx; y_pdc += sy
    cols = {"x": drift_x, "y": drift_y}
    if "z" in locs.columns:
        ref_x = x_pdc[frame <= segmentation]; ref_y = y_pdc[frame <= segmentation]
        ref_z = locs["z"][frame <= segmentation]
        z_pdc, dz1 = intersection_max_z(x_pdc, y_pdc, locs["z"], ref_x, ref_y, ref_z, frame, seg_bounds,
                                        intersect_d, roi_r, width, height, pixelsize, 1, rec3)
        z_pdc, dz2 = intersection_max_z(x_pdc, y_pdc, z_pdc, x_pdc, y_pdc, z_pdc, frame, seg_bounds,
                                        intersect_d, roi_r, width, height, pixelsize, 2, rec3)
        drift_z = dz1 + dz2
        sz = np.mean(drift_z)
        drift_z -= sz
        z_pdc += sz
        cols["z"] = drift_z
        locs["z"] = z_pdc
    locs["x"] = x_pdc
    locs["y"] = y_pdc
    if record is not None:
        record["roi_cc"] = rec2
        record["roi_cc_z"] = rec3
    return locs, pd.DataFrame(cols, dtype="float32")
