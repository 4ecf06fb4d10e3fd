"""Astigmatic 3-D z fitting on B200.

Drop-in for the fitting part of ``picasso.zfit`` (reference picasso/zfit.py): ``zfit`` :465,
``fit_z`` :294, ``fit_z_parallel`` :385, ``filter_z_fits`` :675,
``axial_localization_precision`` :706 and ``axial_localization_precision_astig`` :747 keep their
signatures, returned DataFrames and error behaviour.  The per-localization
``scipy.optimize.minimize_scalar`` loop (:338-355) and the z / d_zcalib / lpz column arithmetic
run in one CUDA kernel (csrc/zfit.cu) through the C ABI (``pb_zfit``).  For tables of 4-byte columns
(everything the fit path produces) ``_fit_z`` is ONE fused device call (``pb_locs_filter``,
csrc/table.cu): upload the columns once, z fit on the resident columns, ``lib.ensure_sanity`` as a
fused mask + stream compaction, the RMSD filter with numpy's float32 pairwise sum reproduced bit for
bit, one download of the kept rows.  ``calibrate_z`` (the calibration itself) is outside the hot path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Literal

import numpy as np
import pandas as pd

from . import __version__, _lib, lib


def _declare(l):
    if getattr(l, "_zfit_declared", False):
        return
    vp, f64, i32, sz = C.c_void_p, C.c_double, C.c_int, C.c_size_t
    l.pb_zfit.argtypes = [sz, vp, vp, vp, vp, vp, vp, vp, vp, f64, f64, i32, vp, vp, vp, vp]
    l.pb_zfit.restype = i32
    l._zfit_declared = True


def _method_id(fitting_method, columns) -> int:
    if fitting_method == "gausslq":
        return 0
    if fitting_method == "gaussmle":
        return 2 if ("sx_unc" in columns and "sy_unc" in columns) else 1
    raise ValueError("fitting_method must be 'gausslq' or 'gaussmle'.")


def _run(locs, cx, cy, magnification_factor, pixelsize, fitting_method, want_nfev=False):
    """(z, d_zcalib, lpz[, nfev]) float32 columns for ``locs`` from the GPU."""
    l = _lib.load()
    _declare(l)
    _lib.require_gpu()
    method = _method_id(fitting_method, locs.columns)
    f32 = lambda name: np.ascontiguousarray(locs[name].to_numpy(), dtype=np.float32)   # noqa: E731
    sx, sy, ph, bg = f32("sx"), f32("sy"), f32("photons"), f32("bg")
    sxu = f32("sx_unc") if method == 2 else None
    syu = f32("sy_unc") if method == 2 else None
    cx = np.ascontiguousarray(cx, dtype=np.float64)
    cy = np.ascontiguousarray(cy, dtype=np.float64)
    if cx.shape != (7,) or cy.shape != (7,):
        raise ValueError("calibration needs 7 X and 7 Y coefficients")
    n = len(sx)
    z = np.zeros(n, np.float32)
    dz = np.zeros(n, np.float32)
    lpz = np.zeros(n, np.float32)
    nfev = np.zeros(n, np.int32) if want_nfev else None
    p = lambda a: _lib.ptr(a) if a is not None else None   # noqa: E731
    _lib.check(l.pb_zfit(n, p(sx), p(sy), p(ph), p(bg), p(sxu), p(syu), p(cx), p(cy),
                         float(magnification_factor), float(pixelsize), method, p(z), p(dz), p(lpz),
                         p(nfev)))
    return (z, dz, lpz, nfev) if want_nfev else (z, dz, lpz)


def filter_z_fits(locs: pd.DataFrame, range: int) -> pd.DataFrame:
    """Drop fits whose calibration residual exceeds ``range`` x RMSD (reference zfit.py:675-704)."""
    if "d_zcalib" not in locs.columns:
        return locs
    if range > 0:
        rmsd = np.sqrt(np.nanmean(locs["d_zcalib"] ** 2))
        locs = locs[locs["d_zcalib"] <= range * rmsd]
    return locs


def _progress(n, progress_callback):
    if progress_callback == "console":
        from tqdm import tqdm

        for _ in tqdm(range(n), desc="Fitting z...", unit="locs"):
            pass
    elif callable(progress_callback):
        for i in range(n):
            progress_callback(i)


def _fit_z(locs, info, calibration, magnification_factor, pixelsize, fitting_method="gausslq",
           filter=2, progress_callback=None):
    """Reference ``_fit_z`` (zfit.py:327-383): z fit, ``locs["z" | "d_zcalib" | "lpz"]`` assigned,
    ``lib.ensure_sanity``, ``filter_z_fits``.  Tables of 4-byte columns take the fused device path."""
    cx = np.array(calibration["X Coefficients"])
    cy = np.array(calibration["Y Coefficients"])
    new_cols = ("z", "d_zcalib", "lpz")
    base = locs[[c for c in locs.columns if c not in new_cols]] if any(c in locs.columns for c in new_cols) else locs
    if len(locs) and lib._device_table_ok(base):
        lib._check_sanity_keys(info)
        method = _method_id(fitting_method, base.columns)
        if cx.shape != (7,) or cy.shape != (7,):
            raise ValueError("calibration needs 7 X and 7 Y coefficients")
        names = list(base.columns)
        ci = lambda c: names.index(c) if c in names else -1      # noqa: E731
        spec = lib.ZfitSpec(ci("sx"), ci("sy"), ci("photons"), ci("bg"),
                            ci("sx_unc") if method == 2 else -1, ci("sy_unc") if method == 2 else -1,
                            (C.c_double * 7)(*cx.astype(np.float64)), (C.c_double * 7)(*cy.astype(np.float64)),
                            float(magnification_factor), float(pixelsize), method, int(filter))
        out = lib._filter_table(base, info, zspec=spec, extra_names=new_cols)
        _progress(len(locs), progress_callback)
        # column order of the reference: existing z / d_zcalib / lpz columns keep their place
        order = list(locs.columns) + [c for c in new_cols if c not in locs.columns]
        return out if list(out.columns) == order else out[order]
    locs = locs.copy()
    z, dz, lpz = _run(locs, cx, cy, magnification_factor, pixelsize, fitting_method)
    _progress(len(z), progress_callback)
    locs["z"] = z
    locs["d_zcalib"] = dz
    locs["lpz"] = lpz
    locs = lib.ensure_sanity(locs, info)
    return filter_z_fits(locs, filter)


def fit_z(locs, info, calibration, magnification_factor, pixelsize,
          fitting_method: Literal["gausslq", "gaussmle"] = "gausslq", filter: int = 2,
          progress_callback=None) -> pd.DataFrame:
    """Reference ``fit_z`` (zfit.py:294-324; deprecated upstream in favour of ``zfit``)."""
    return _fit_z(locs, info, calibration, magnification_factor, pixelsize, fitting_method, filter,
                  progress_callback)


def fit_z_parallel(locs, info, calibration, magnification_factor, pixelsize,
                   fitting_method: Literal["gausslq", "gaussmle"] = "gausslq", filter: int = 2,
                   asynch: bool = False):
    """Reference ``fit_z_parallel`` (zfit.py:385-413).  The GPU fits all localizations in one
    launch; with ``asynch=True`` a list with one completed future is returned so that
    ``locs_from_futures`` applies unchanged."""
    if asynch:
        from concurrent.futures import Future

        f = Future()
        f.set_result(_fit_z(locs, info, calibration, magnification_factor, pixelsize, fitting_method, 0))
        return [f]
    return _fit_z(locs, info, calibration, magnification_factor, pixelsize, fitting_method, filter)


def locs_from_futures(futures, filter: int = 2) -> pd.DataFrame:
    """Reference ``locs_from_futures`` (zfit.py:648-672)."""
    locs = pd.concat([f.result() for f in futures], ignore_index=True)
    return filter_z_fits(locs, filter)


def zfit(locs: pd.DataFrame, info: list[dict], *, calibration: dict,
         magnification_factor: float | None = None, pixelsize: int | float | None = None,
         fitting_method: Literal["gausslq", "gaussmle"] = "gausslq", filter: int = 2,
         multiprocess: bool = False,
         progress_callback: Callable[[int], None] | Literal["console"] | None = None,
         abort_callback: Callable[[], bool] | None = None):
    """Fit z coordinates (reference ``zfit``, zfit.py:465-646): returns ``(locs, info)`` with
    columns ``z``, ``d_zcalib``, ``lpz`` appended, or ``(None, None)`` when aborted."""
    assert fitting_method in ["gausslq", "gaussmle"], "Invalid fitting method."
    assert filter >= 0, "Filter must be non-negative."
    assert isinstance(calibration, dict), "Calibration must be a dict, see ``io.load_calibration``."
    if magnification_factor is not None:
        assert isinstance(magnification_factor, (int, float)), "Magnification factor must be a number."
        calibration["Magnification factor"] = float(magnification_factor)
    else:
        assert "Magnification factor" in calibration, "Magnification factor is missing in calibration."
    if pixelsize is not None:
        assert isinstance(pixelsize, (int, float)), "Pixelsize must be a number in nm."
        pixelsize = float(pixelsize)
        info.append({"Pixelsize": pixelsize})
    else:
        assert lib.get_from_metadata(info, "Pixelsize") is not None, (
            "Camera pixel size (nm) is missing. Enter it either in the info metadata, or as an "
            "argument.")
    pixelsize = lib.get_from_metadata(info, "Pixelsize", raise_error=True)
    if abort_callback is not None and abort_callback():
        return None, None
    if multiprocess and callable(progress_callback):
        progress_callback(len(locs))
        progress_callback = None
    locs = _fit_z(locs, info, calibration, calibration["Magnification factor"], pixelsize,
                  fitting_method=fitting_method, filter=filter, progress_callback=progress_callback)
    new_info = {
        "Generated by": f"Picasso v{__version__} Fit 3D (picasso_b200)",
        "Calibration path": calibration.get("Path", "N/A"),
        "Filter range": filter,
    }
    return locs, info + [new_info | calibration]


def axial_localization_precision_astig(locs, info, calibration,
                                       fitting_method: Literal["gausslq", "gaussmle"] = "gausslq"):
    """Axial localization precision for fitted localizations (reference zfit.py:747-802); the
    column arithmetic of ``_axial_localization_precision_astig`` (:805-890) on the GPU.  ``locs``
    must contain ``z`` (nm); returns lpz in nm."""
    assert fitting_method in ["gausslq", "gaussmle"], "fitting_method must be 'gausslq' or 'gaussmle'."
    assert ("X Coefficients" in calibration and "Y Coefficients" in calibration
            and "Magnification factor" in calibration), (
        "Calibration dictionary must contain 'X Coefficients', 'Y Coefficients', and "
        "'Magnification factor'.")
    pixelsize = lib.get_from_metadata(info, "Pixelsize")
    if pixelsize is None:
        raise ValueError("Pixelsize not found in info.")
    locs = pd.DataFrame(locs)
    # the kernel derives lpz from its own z; re-fit is avoided by evaluating the same float32
    # expressions on the host for a given z column
    return _lpz_host(locs, np.array(calibration["X Coefficients"]), np.array(calibration["Y Coefficients"]),
                     calibration["Magnification factor"], pixelsize, fitting_method)


def axial_localization_precision(locs, info, calibration,
                                 fitting_method: Literal["gausslq", "gaussmle"] = "gausslq",
                                 modality: Literal["astigmatic"] = "astigmatic"):
    """Reference zfit.py:706-744."""
    if modality != "astigmatic":
        raise NotImplementedError("Currently only 'astigmatic' modality is supported.")
    return axial_localization_precision_astig(locs, info, calibration, fitting_method)


def _lpz_host(locs, cx, cy, magnification_factor, pixelsize, fitting_method):
    """``_axial_localization_precision_astig`` (zfit.py:805-890) for an existing ``z`` column."""
    from . import gausslq, gaussmle

    if fitting_method == "gausslq":
        se_sx = gausslq.sigma_uncertainty(locs["sx"], locs["sy"], locs["photons"], locs["bg"]) * pixelsize
        se_sy = gausslq.sigma_uncertainty(locs["sy"], locs["sx"], locs["photons"], locs["bg"]) * pixelsize
    elif fitting_method == "gaussmle":
        if "sx_unc" not in locs.columns or "sy_unc" not in locs.columns:
            se_sx = gaussmle.sigma_uncertainty(locs["sx"], locs["sy"], locs["photons"], locs["bg"]) * pixelsize
            se_sy = gaussmle.sigma_uncertainty(locs["sy"], locs["sx"], locs["photons"], locs["bg"]) * pixelsize
        else:
            se_sx = locs["sx_unc"] * pixelsize
            se_sy = locs["sy_unc"] * pixelsize
    else:
        raise ValueError("fitting_method must be 'gausslq' or 'gaussmle'.")
    z = locs["z"] / magnification_factor

    def size(c):
        return c[0] * z**6 + c[1] * z**5 + c[2] * z**4 + c[3] * z**3 + c[4] * z**2 + c[5] * z + c[6]

    def prime(c):
        return 6 * c[0] * z**5 + 5 * c[1] * z**4 + 4 * c[2] * z**3 + 3 * c[3] * z**2 + 2 * c[4] * z + c[5]

    wx, wy = size(cx) * pixelsize, size(cy) * pixelsize
    wxp, wyp = prime(cx) * pixelsize, prime(cy) * pixelsize
    swx, swy = np.sqrt(wx), np.sqrt(wy)
    a2, b2 = (wxp / (2 * swx)) ** 2, (wyp / (2 * swy)) ** 2
    c2 = ((1 / (2 * np.sqrt(locs["sx"] * pixelsize))) * se_sx) ** 2
    d2 = ((1 / (2 * np.sqrt(locs["sy"] * pixelsize))) * se_sy) ** 2
    return np.sqrt((a2 * c2 + b2 * d2) / (a2 + b2) ** 2) * magnification_factor
