"""Spot identification, ROI extraction and 2-D fitting glue on B200.

Drop-in for the hot-path part of ``picasso.localize`` (reference
picasso/localize.py): ``identify_in_image`` :247, ``identify_in_frame`` :295,
``identify_by_frame_number`` :340, ``identify`` :639, ``get_spots`` :1115,
``fit2D`` :1344 and ``localize`` :1682 keep their signatures, return types and
error behaviour; the numba kernels underneath are replaced by CUDA
(csrc/identify.cu, csrc/mle_fit.cu, csrc/lq_fit.cu) through the C ABI.

Movies are streamed through the GPU in frame chunks, so ``np.memmap`` movies
(the reference's lazy .raw reader, io.py:50-96) are read from disk exactly once
per pass.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Callable, Literal

import numpy as np
import pandas as pd

from . import __version__, _lib

_CHUNK_BYTES = 256 << 20


def _declare(lib):
    if getattr(lib, "_localize_declared", False):
        return
    vp, i32, i64, f64, f32, sz = (C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_float,
                                  C.c_size_t)
    lib.pb_identify.argtypes = [vp, i32, sz, i32, i32, i64, i32, f64, vp, vp, vp, vp, vp, sz,
                                C.POINTER(sz)]
    lib.pb_identify.restype = i32
    lib.pb_get_spots.argtypes = [vp, i32, sz, i32, i32, i64, sz, vp, vp, vp, i32, f32, f32, f32, vp]
    lib.pb_get_spots.restype = i32
    lib.pb_identify_get_spots.argtypes = [vp, i32, sz, i32, i32, i64, i32, f64, vp, f32, f32, f32,
                                          vp, vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.pb_identify_get_spots.restype = i32
    lib.pb_locs_columns.argtypes = [i32]
    lib.pb_locs_columns.restype = i32
    lib.pb_localize.argtypes = [vp, i32, sz, i32, i32, i64, i32, f64, vp, f32, f32, f32, i32, f64, i32,
                                i32, vp, sz, C.POINTER(sz)]
    lib.pb_localize.restype = i32
    lib.pb_locs_from_fits.argtypes = [sz, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.pb_locs_from_fits.restype = i32
    lib._localize_declared = True


def _lib_ready():
    lib = _lib.load()
    _declare(lib)
    _lib.require_gpu()
    return lib


def _as_device_movie(chunk):
    """uint16 movies go to the GPU as they are; anything else as float32
    (the reference casts every frame with np.float32(frame), localize.py:332)."""
    chunk = np.asarray(chunk)
    if chunk.dtype == np.uint16:
        return np.ascontiguousarray(chunk), 0
    return np.ascontiguousarray(chunk, dtype=np.float32), 1


def _roi_array(roi):
    if roi is None:
        return None
    return np.array([roi[0][0], roi[0][1], roi[1][0], roi[1][1]], dtype=np.int32)


def _identify_chunk(lib, chunk, frame_offset, minimum_ng, box, roi_arr, capacity=None):
    chunk, dtype = _as_device_movie(chunk)
    if chunk.ndim == 2:
        chunk = chunk[None]
    F, Y, X = chunk.shape
    if capacity is None:
        capacity = max(4096, 512 * F)
    while True:
        fr = np.empty(capacity, np.int64)
        xs = np.empty(capacity, np.int64)
        ys = np.empty(capacity, np.int64)
        ng = np.empty(capacity, np.float32)
        found = C.c_size_t(0)
        rc = lib.pb_identify(_lib.ptr(chunk), dtype, F, Y, X, int(frame_offset), int(box),
                             float(minimum_ng), _lib.ptr(roi_arr) if roi_arr is not None else None,
                             _lib.ptr(fr), _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(ng), capacity,
                             C.byref(found))
        if rc == 4:  # PB_ERR_CAPACITY: retry with the size the kernel reported
            capacity = int(found.value)
            continue
        _lib.check(rc)
        n = int(found.value)
        return fr[:n].copy(), xs[:n].copy(), ys[:n].copy(), ng[:n].copy()


def identify_in_image(image, minimum_ng: float, box: int):
    """Local maxima with net gradient above ``minimum_ng`` in one image.

    Same as reference ``identify_in_image`` (localize.py:247-292): returns
    ``(y, x, net_gradient)`` in row-major order.
    """
    lib = _lib_ready()
    image = np.ascontiguousarray(image, dtype=np.float32)
    _, xs, ys, ng = _identify_chunk(lib, image, 0, minimum_ng, box, None)
    return ys, xs, ng


def identify_in_frame(frame, minimum_ng: float, box: int, roi=None):
    """Reference ``identify_in_frame`` (localize.py:295-337): optional ROI
    ``((y0, x0), (y1, x1))`` is cut out first; coordinates refer to the full frame."""
    lib = _lib_ready()
    _, xs, ys, ng = _identify_chunk(lib, np.asarray(frame), 0, minimum_ng, box, _roi_array(roi))
    return ys, xs, ng


def _empty_ids():
    return pd.DataFrame(
        {
            "frame": pd.Series(dtype=int),
            "x": pd.Series(dtype=int),
            "y": pd.Series(dtype=int),
            "net_gradient": pd.Series(dtype=np.float32),
        }
    )


def _frame_range(n_frames, frame_bounds):
    """Inclusive [lo, hi] frame numbers kept by the reference's bounds check
    (localize.py:395-409)."""
    lo, hi = 0, n_frames
    if frame_bounds is not None:
        if frame_bounds[0] is not None:
            lo = max(frame_bounds[0], lo)
        if frame_bounds[1] is not None:
            hi = min(frame_bounds[1], hi)
    return lo, hi


def identify_by_frame_number(movie, minimum_ng: float, box: int, frame_number: int, *, roi=None,
                             frame_bounds=None, lock=None) -> pd.DataFrame:
    """Reference ``identify_by_frame_number`` (localize.py:340-421)."""
    if lock is not None:
        with lock:
            frame = movie[frame_number]
    else:
        frame = movie[frame_number]
    if frame_bounds is not None:
        lo, hi = _frame_range(len(movie), frame_bounds)
        if not (lo <= frame_number <= hi):
            return _empty_ids()
    y, x, ng = identify_in_frame(frame, minimum_ng, box, roi)
    return pd.DataFrame(
        {
            "frame": np.full(len(x), frame_number, dtype=int),
            "x": x.astype(int),
            "y": y.astype(int),
            "net_gradient": ng.astype(np.float32),
        }
    )


def _movie_chunk(movie, f0, f1):
    try:
        return np.asarray(movie[f0:f1])
    except (TypeError, IndexError, ValueError):
        return np.stack([np.asarray(movie[i]) for i in range(f0, f1)])


def _frames_per_chunk(movie):
    first = np.asarray(movie[0])
    per_frame = max(1, first.size * max(2, first.dtype.itemsize))
    return max(1, _CHUNK_BYTES // per_frame)


def identify(movie, minimum_ng: float, box: int, *, roi=None, frame_bounds=None,
             threaded: bool = True, progress_callback=None, abort_callback=None,
             return_info: bool = None):
    """Identify spots in every frame of ``movie``.

    Reference ``identify`` (localize.py:639-749): returns a DataFrame
    ``{frame:int64, x:int64, y:int64, net_gradient:float32}`` sorted by frame
    (and the metadata dict when ``return_info=True``); ``None`` if aborted.
    ``threaded`` only changes progress reporting here -- frames are processed by
    the GPU in chunks either way.
    """
    if return_info is None:
        return_info = False
        warnings.warn(
            "Warning: In Picasso v0.11.0, picasso.localize.identify() will return both the "
            "identifications and a metadata dictionary by default. Please pass 'return_info' "
            "explicitly.", DeprecationWarning, stacklevel=2)
    lib = _lib_ready()
    N = len(movie)
    lo, hi = _frame_range(N, frame_bounds)
    roi_arr = _roi_array(roi)
    use_tqdm = progress_callback == "console"
    bar = None
    if use_tqdm:
        from tqdm import tqdm

        bar = tqdm(total=N, desc="Identifying spots", unit="frame")
    parts = []
    step = _frames_per_chunk(movie) if N else 1
    f = 0
    while f < N:
        if abort_callback is not None and abort_callback():
            if bar is not None:
                bar.close()
            return None
        f1 = min(N, f + step)
        a, b = max(f, lo), min(f1, hi + 1)      # frames of this chunk inside the bounds
        if a < b:
            parts.append(_identify_chunk(lib, _movie_chunk(movie, a, b), a, minimum_ng, box,
                                         roi_arr))
        if bar is not None:
            bar.update(f1 - f)
        elif callable(progress_callback):
            if threaded:
                progress_callback(f1)
            else:
                for i in range(f, f1):
                    progress_callback(i)
        f = f1
    if bar is not None:
        bar.close()
    if parts:
        fr = np.concatenate([p[0] for p in parts])
        xs = np.concatenate([p[1] for p in parts])
        ys = np.concatenate([p[2] for p in parts])
        ng = np.concatenate([p[3] for p in parts])
    else:
        fr = xs = ys = np.zeros(0, np.int64)
        ng = np.zeros(0, np.float32)
    ids = pd.DataFrame({"frame": fr.astype(int), "x": xs.astype(int), "y": ys.astype(int),
                        "net_gradient": ng.astype(np.float32)})
    if return_info:
        info = {
            "Generated by": f"Picasso: v{__version__} Identify (picasso_b200)",
            "Min. Net Gradient": minimum_ng,
            "Box Size": box,
            "ROI": roi,
            "Frame Bounds": frame_bounds,
        }
        return ids, info
    return ids


def identify_async(movie, minimum_ng: float, box: int, *, roi=None, frame_bounds=None):
    """Start identification in the background (reference ``identify_async``,
    localize.py:482-558): returns ``(current, futures)`` immediately; ``current[0]`` counts
    the frames processed so far and reaches ``len(movie)`` when done;
    ``identifications_from_futures(futures)`` gives the final table.  The reference fans
    frames out to a thread pool; here one worker thread streams frame chunks through the
    GPU."""
    from concurrent.futures import ThreadPoolExecutor

    lib = _lib_ready()
    N = len(movie)
    lo, hi = _frame_range(N, frame_bounds)
    roi_arr = _roi_array(roi)
    current = [0]

    def _work():
        out = []
        step = _frames_per_chunk(movie) if N else 1
        f = 0
        while f < N:
            f1 = min(N, f + step)
            a, b = max(f, lo), min(f1, hi + 1)
            if a < b:
                fr, xs, ys, ng = _identify_chunk(lib, _movie_chunk(movie, a, b), a, minimum_ng, box,
                                                 roi_arr)
                out.append(pd.DataFrame({"frame": fr.astype(int), "x": xs.astype(int),
                                         "y": ys.astype(int),
                                         "net_gradient": ng.astype(np.float32)}))
            current[0] = f1
            f = f1
        if not out:
            out.append(_empty_ids())
        return out

    executor = ThreadPoolExecutor(1)
    futures = [executor.submit(_lib.on_callers_device(_work))]
    executor.shutdown(wait=False)
    return current, futures


def identifications_from_futures(futures) -> pd.DataFrame:
    """Combine the results of ``identify_async`` (reference localize.py:457-479)."""
    from itertools import chain

    ids_list = list(chain(*[f.result() for f in futures]))
    ids = pd.concat(ids_list, ignore_index=True)
    ids.sort_values(by="frame", kind="quicksort", inplace=True)
    return ids


def get_spots(movie, identifications: pd.DataFrame, box: int, camera_info: dict):
    """Cut ``box x box`` ROIs around the identifications and convert to photons.

    Reference ``get_spots`` (localize.py:1115-1145): returns float32
    ``(k, box, box)`` = ``(spots - Baseline) * Sensitivity / Gain``.
    """
    lib = _lib_ready()
    fr = np.ascontiguousarray(identifications["frame"].to_numpy(), dtype=np.int64)
    xs = np.ascontiguousarray(identifications["x"].to_numpy(), dtype=np.int64)
    ys = np.ascontiguousarray(identifications["y"].to_numpy(), dtype=np.int64)
    n = len(fr)
    spots = np.zeros((n, box, box), dtype=np.float32)
    if n == 0:
        return spots
    baseline = float(camera_info["Baseline"])
    sensitivity = float(camera_info["Sensitivity"])
    gain = float(camera_info["Gain"])
    N = len(movie)
    step = _frames_per_chunk(movie)
    fmin, fmax = int(fr.min()), int(fr.max())
    if fmin < 0 or fmax >= N:
        raise IndexError("identification frame out of range for movie")
    for f0 in range(fmin, fmax + 1, step):
        f1 = min(fmax + 1, f0 + step)
        if not np.any((fr >= f0) & (fr < f1)):
            continue
        chunk, dtype = _as_device_movie(_movie_chunk(movie, f0, f1))
        F, Y, X = chunk.shape
        _lib.check(lib.pb_get_spots(_lib.ptr(chunk), dtype, F, Y, X, f0, n, _lib.ptr(fr),
                                    _lib.ptr(xs), _lib.ptr(ys), int(box), baseline, sensitivity,
                                    gain, _lib.ptr(spots)))
    return spots


def _identify_and_cut(movie, minimum_ng, box, camera_info, roi=None, frame_bounds=None,
                      progress_callback=None):
    """identify + get_spots in one pass over the movie (each frame chunk is uploaded once).
    Returns (identifications DataFrame, spots float32 (n, box, box)); identical to
    ``identify`` followed by ``get_spots``."""
    lib = _lib_ready()
    N = len(movie)
    lo, hi = _frame_range(N, frame_bounds)
    roi_arr = _roi_array(roi)
    baseline = float(camera_info["Baseline"])
    sensitivity = float(camera_info["Sensitivity"])
    gain = float(camera_info["Gain"])
    parts = []
    step = _frames_per_chunk(movie) if N else 1
    f = 0
    while f < N:
        f1 = min(N, f + step)
        a, b = max(f, lo), min(f1, hi + 1)
        if a < b:
            chunk, dtype = _as_device_movie(_movie_chunk(movie, a, b))
            F, Y, X = chunk.shape
            capacity = max(4096, 512 * F)
            while True:
                fr = np.empty(capacity, np.int64); xs = np.empty(capacity, np.int64)
                ys = np.empty(capacity, np.int64); ng = np.empty(capacity, np.float32)
                sp = np.empty((capacity, box, box), np.float32)
                found = C.c_size_t(0)
                rc = lib.pb_identify_get_spots(
                    _lib.ptr(chunk), dtype, F, Y, X, a, int(box), float(minimum_ng),
                    _lib.ptr(roi_arr) if roi_arr is not None else None, baseline, sensitivity, gain,
                    _lib.ptr(fr), _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(ng), _lib.ptr(sp), capacity,
                    C.byref(found))
                if rc == 4:
                    capacity = int(found.value)
                    continue
                _lib.check(rc)
                n = int(found.value)
                parts.append((fr[:n], xs[:n], ys[:n], ng[:n], sp[:n]))
                break
        if callable(progress_callback):
            progress_callback(f1)
        f = f1
    if parts:
        cat = [np.concatenate([p[k] for p in parts]) for k in range(5)]
    else:
        cat = [np.zeros(0, np.int64)] * 3 + [np.zeros(0, np.float32),
                                             np.zeros((0, box, box), np.float32)]
    ids = pd.DataFrame({"frame": cat[0].astype(int), "x": cat[1].astype(int),
                        "y": cat[2].astype(int), "net_gradient": cat[3].astype(np.float32)})
    return ids, np.ascontiguousarray(cat[4])


LOCS_COLUMNS_MLE = ("frame", "x", "y", "photons", "sx", "sy", "bg", "lpx", "lpy", "ellipticity",
                    "net_gradient", "log_likelihood", "iterations", "photons_unc", "bg_unc",
                    "sx_unc", "sy_unc")
LOCS_COLUMNS_LQ = LOCS_COLUMNS_MLE[:11]
_UINT_COLUMNS = ("frame", "iterations")
_FIT_IDS = {("gaussmle", "sigma"): 0, ("gaussmle", "sigmaxy"): 1, ("gausslq", None): 2,
            ("gausslq-gpu", None): 3}
_FUSED_CALL_BYTES = 1 << 30


def _columns_to_locs(cols, names):
    """(ncols, n) 4-byte column block -> DataFrame with the reference's dtypes (zero-copy
    views), then the reference's final ``sort_values(by="frame", kind="quicksort")``."""
    data = {}
    for k, name in enumerate(names):
        data[name] = cols[k].view(np.uint32) if name in _UINT_COLUMNS else cols[k]
    locs = pd.DataFrame(data, copy=False)
    locs.sort_values(by="frame", kind="quicksort", inplace=True)
    return locs


def locs_columns_from_fits(identifications: pd.DataFrame, theta, box: int, fit: int, em: bool = False,
                           CRLBs=None, log_likelihoods=None, iterations=None):
    """The localization-table column arithmetic of ``gaussmle.locs_from_fits`` (fit 0/1),
    ``gausslq.locs_from_fits`` (2) and ``locs_from_fits_gpufit`` (3) evaluated on the GPU
    (csrc/localize.cu): returns ``{column: array}`` in identification order (unsorted)."""
    lib = _lib_ready()
    n = len(identifications)
    names = LOCS_COLUMNS_MLE if fit <= 1 else LOCS_COLUMNS_LQ
    cols = np.zeros((len(names), n), np.float32)
    if n:
        c64 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.int64)   # noqa: E731
        c32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)   # noqa: E731
        fr, xs, ys = (c64(identifications[k]) for k in ("frame", "x", "y"))
        ng = c32(identifications["net_gradient"])
        th = c32(theta)
        cr = c32(CRLBs) if fit <= 1 else None
        ll = c32(log_likelihoods) if fit <= 1 else None
        it = np.ascontiguousarray(np.asarray(iterations), dtype=np.int32) if fit <= 1 else None
        _lib.check(lib.pb_locs_from_fits(n, int(fit), int(box), int(bool(em)), _lib.ptr(fr), _lib.ptr(xs),
                                         _lib.ptr(ys), _lib.ptr(ng), _lib.ptr(th),
                                         _lib.ptr(cr) if cr is not None else None,
                                         _lib.ptr(ll) if ll is not None else None,
                                         _lib.ptr(it) if it is not None else None, _lib.ptr(cols)))
    return {name: (cols[k].view(np.uint32) if name in _UINT_COLUMNS else cols[k])
            for k, name in enumerate(names)}


def _localize_fused(movie, minimum_ng, box, camera_info, fit, eps, max_it, roi=None,
                    frame_bounds=None, progress_callback=None):
    """movie -> localization table without leaving the GPU in between (``pb_localize``):
    identify -> get_spots -> fit -> locs_from_fits per frame chunk; only the finished columns
    come back.  Same result as ``identify`` + ``fit2D``."""
    names = LOCS_COLUMNS_MLE if fit <= 1 else LOCS_COLUMNS_LQ
    cols = _localize_fused_columns(movie, minimum_ng, box, camera_info, fit, eps, max_it, roi=roi,
                                   frame_bounds=frame_bounds, progress_callback=progress_callback)
    return _columns_to_locs(cols, names)


def _localize_fused_columns(movie, minimum_ng, box, camera_info, fit, eps, max_it, roi=None,
                            frame_bounds=None, progress_callback=None):
    """The (ncols, n) float32 column block of ``_localize_fused`` in (frame, y, x) order."""
    lib = _lib_ready()
    N = len(movie)
    lo, hi = _frame_range(N, frame_bounds)
    roi_arr = _roi_array(roi)
    baseline = float(camera_info["Baseline"])
    sensitivity = float(camera_info["Sensitivity"])
    gain = float(camera_info["Gain"])
    em = int(camera_info["Gain"] > 1)
    names = LOCS_COLUMNS_MLE if fit <= 1 else LOCS_COLUMNS_LQ
    assert lib.pb_locs_columns(fit) == len(names)
    parts = []
    if N:
        first = np.asarray(movie[0])
        per_frame = max(1, first.size * (2 if first.dtype == np.uint16 else 4))
        step = max(1, _FUSED_CALL_BYTES // per_frame)
    f = 0
    while f < N:
        f1 = min(N, f + step)
        a, b = max(f, lo), min(f1, hi + 1)
        if a < b:
            chunk, dtype = _as_device_movie(_movie_chunk(movie, a, b))
            F, Y, X = chunk.shape
            capacity = max(4096, 128 * F)
            while True:
                cols = _lib.pinned_empty((len(names), capacity), np.float32)
                found = C.c_size_t(0)
                rc = lib.pb_localize(_lib.ptr(chunk), dtype, F, Y, X, a, int(box), float(minimum_ng),
                                     _lib.ptr(roi_arr) if roi_arr is not None else None, baseline,
                                     sensitivity, gain, int(fit), float(eps), int(max_it), em,
                                     _lib.ptr(cols), capacity, C.byref(found))
                if rc == 4:
                    capacity = int(found.value)
                    continue
                _lib.check(rc)
                parts.append(cols[:, :int(found.value)])
                break
        if callable(progress_callback):
            progress_callback(f1)
        f = f1
    if len(parts) == 1:
        cols = parts[0]
    elif parts:
        cols = np.concatenate(parts, axis=1)
    else:
        cols = np.zeros((len(names), 0), np.float32)
    return cols


def _fit_spots(spots, identifications, box, camera_info, fitting_method, eps, max_it, mle_method,
               progress_callback):
    from . import gausslq, gaussmle

    em = camera_info["Gain"] > 1
    if fitting_method == "gausslq":
        theta = gausslq.fit_spots(spots, progress_callback)
        return gausslq.locs_from_fits(identifications, theta, box, em)
    if fitting_method == "gausslq-gpu":
        if callable(progress_callback):
            progress_callback(1)
        theta = gausslq.fit_spots_gpufit(spots)
        return gausslq.locs_from_fits_gpufit(identifications, theta, box, em)
    if fitting_method == "gaussmle":
        thetas, CRLBs, llhoods, iterations = gaussmle.gaussmle(spots, eps, max_it, mle_method,
                                                               progress_callback)
        return gaussmle.locs_from_fits(identifications, thetas, CRLBs, llhoods, iterations, box)
    raise NotImplementedError(
        "fitting_method='avg' (picasso.avgroi) is outside the B200 hot path; "
        "use the reference implementation")


def fit2D(movie, movie_info, camera_info, identifications, box,
          fitting_method: Literal["gausslq", "gausslq-gpu", "gaussmle", "avg"] = "gausslq",
          eps: float = 0.001, max_it: int = 100,
          mle_method: Literal["sigma", "sigmaxy"] = "sigmaxy", multiprocess: bool = True,
          progress_callback=None, abort_callback=None):
    """Fit 2-D localizations (reference ``fit2D``, localize.py:1344-1506).

    Returns ``(locs DataFrame | None, new_info dict)``.  ``multiprocess`` is
    accepted for compatibility; every method runs on the GPU.
    """
    assert hasattr(movie, "__getitem__") and hasattr(movie, "__len__"), \
        "movie must be a movie loaded by picasso.io.load_movie"
    assert isinstance(movie_info, list), "movie_info must be a list"
    assert isinstance(camera_info, dict), "camera_info must be a dict"
    assert isinstance(identifications, pd.DataFrame), "identifications must be a DataFrame"
    assert isinstance(box, int) and box > 0, "box must be a positive integer"
    assert fitting_method in ["gausslq", "gausslq-gpu", "gaussmle", "avg"], (
        "fitting_method must be one of 'gausslq', 'gausslq-gpu', 'gaussmle', or 'avg'")
    assert isinstance(eps, (int, float)) and eps > 0, "eps must be a positive number"
    assert isinstance(max_it, int) and max_it > 0, "max_it must be a positive integer"
    assert mle_method in ["sigma", "sigmaxy"], "mle_method must be 'sigma' or 'sigmaxy'"
    assert isinstance(multiprocess, bool), "multiprocess must be a boolean"
    if "Pixelsize" not in camera_info:
        warnings.warn("Camera info in picasso.localize.fit2D does not contain 'Pixelsize', "
                      "i.e., effective camera pixel size in nm. Assuming 130.")
        camera_info["Pixelsize"] = 130

    if callable(abort_callback) and abort_callback():
        locs = None
    else:
        spots = get_spots(movie, identifications, box, camera_info)
        locs = _fit_spots(spots, identifications, box, camera_info, fitting_method, eps, max_it,
                          mle_method, progress_callback)
    localize_info = {
        "Generated by": f"Picasso: v{__version__} Fit 2D (picasso_b200)",
        "Fit method": fitting_method,
    }
    if fitting_method == "gaussmle":
        localize_info["Convergence criterion"] = eps
        localize_info["Max iterations"] = max_it
    new_info = localize_info | camera_info
    return locs, new_info


def localize(movie, camera_info: dict, parameters: dict, *, roi=None, frame_bounds=None,
             movie_info=None,
             fitting_method: Literal["gausslq", "gausslq-gpu", "gaussmle", "avg"] = "gausslq",
             eps: float = 0.001, max_it: int = 100,
             mle_method: Literal["sigma", "sigmaxy"] = "sigmaxy", threaded: bool = True,
             identification_progress_callback=None, fit_progress_callback=None,
             return_info: bool = None):
    """Identify and fit spots (reference ``localize``, localize.py:1682-1815)."""
    if return_info is None:
        return_info = False
        warnings.warn(
            "Warning: In Picasso v0.11.0, picasso.localize.localize() will return both the "
            "localizations and a metadata dictionary by default. Please pass 'return_info' "
            "explicitly.", DeprecationWarning, stacklevel=2)
    if movie_info is None:
        movie_info = []
    box = parameters["Box Size"]
    minimum_ng = parameters["Min. Net Gradient"]
    assert isinstance(camera_info, dict), "camera_info must be a dict"
    assert fitting_method in ["gausslq", "gausslq-gpu", "gaussmle", "avg"], (
        "fitting_method must be one of 'gausslq', 'gausslq-gpu', 'gaussmle', or 'avg'")
    if "Pixelsize" not in camera_info:
        warnings.warn("Camera info in picasso.localize.fit2D does not contain 'Pixelsize', "
                      "i.e., effective camera pixel size in nm. Assuming 130.")
        camera_info["Pixelsize"] = 130
    id_cb = identification_progress_callback if callable(identification_progress_callback) else None
    fit_id = _FIT_IDS.get((fitting_method, mle_method if fitting_method == "gaussmle" else None))
    if fit_id is not None and isinstance(box, int) and 5 <= box <= 15 and box % 2 == 1:
        # movie -> locs in one pass: each chunk crosses PCIe once, only the table comes back
        locs = _localize_fused(movie, minimum_ng, box, camera_info, fit_id, eps, max_it, roi=roi,
                               frame_bounds=frame_bounds, progress_callback=id_cb)
        if callable(fit_progress_callback):
            if fitting_method == "gausslq-gpu":
                fit_progress_callback(1)
            else:
                for i in range(len(locs)):
                    fit_progress_callback(i)
    else:
        # identify + get_spots fused (one upload), fit from host ROIs
        identifications, spots = _identify_and_cut(
            movie, minimum_ng, box, camera_info, roi=roi, frame_bounds=frame_bounds,
            progress_callback=id_cb)
        locs = _fit_spots(spots, identifications, box, camera_info, fitting_method, eps, max_it,
                          mle_method, fit_progress_callback)
    identify_info = {
        "Generated by": f"Picasso: v{__version__} Identify (picasso_b200)",
        "Min. Net Gradient": minimum_ng, "Box Size": box, "ROI": roi, "Frame Bounds": frame_bounds,
    }
    fit_info = {"Generated by": f"Picasso: v{__version__} Fit 2D (picasso_b200)",
                "Fit method": fitting_method}
    if fitting_method == "gaussmle":
        fit_info["Convergence criterion"] = eps
        fit_info["Max iterations"] = max_it
    fit_info = fit_info | camera_info
    info = movie_info + [identify_info] + [fit_info]
    if return_info:
        return locs, info
    return locs


def localize_3D(movie, *, movie_info: list, camera_info: dict, box: int, minimum_ng: float,
                calibration_3d: dict, roi=None, frame_bounds=None,
                fitting_method: Literal["gausslq", "gausslq-gpu", "gaussmle"] = "gausslq",
                eps: float = 0.001, max_it: int = 100,
                mle_method: Literal["sigma", "sigmaxy"] = "sigmaxy", multiprocess: bool = True,
                identification_progress_callback=None, fit_progress_callback=None,
                fit_z_progress_callback=None):
    """Identify, fit in 2-D and fit z from astigmatism (reference ``localize_3D``,
    localize.py:1818-2034): the fused movie -> table pass followed by the z-fit kernel.
    Returns ``(locs, info)`` with ``z``, ``d_zcalib``, ``lpz`` appended (no RMSD filter)."""
    from . import zfit

    assert hasattr(movie, "__getitem__") and hasattr(movie, "__len__"), \
        "movie must be a numpy array or ND2Movie"
    assert isinstance(movie_info, list), "movie_info must be a list"
    assert isinstance(camera_info, dict), "camera_info must be a dict"
    assert isinstance(box, int) and box > 0 and box % 2 == 1, "box must be a positive odd integer"
    assert isinstance(minimum_ng, (int, float)), "minimum_ng must be a number"
    assert isinstance(calibration_3d, (dict, str)), \
        "calibration_3d must be a dict or a path to a YAML file"
    assert fitting_method in ["gausslq", "gausslq-gpu", "gaussmle"], \
        "fitting_method must be one of 'gausslq', 'gausslq-gpu', or 'gaussmle'"
    assert isinstance(eps, (int, float)) and eps > 0, "eps must be a positive number"
    assert isinstance(max_it, int) and max_it > 0, "max_it must be a positive integer"
    assert mle_method in ["sigma", "sigmaxy"], "mle_method must be 'sigma' or 'sigmaxy'"
    assert isinstance(multiprocess, bool), "multiprocess must be a boolean"
    locs, info = localize(movie, camera_info, {"Min. Net Gradient": minimum_ng, "Box Size": box},
                          roi=roi, frame_bounds=frame_bounds, movie_info=movie_info,
                          fitting_method=fitting_method, eps=eps, max_it=max_it, mle_method=mle_method,
                          threaded=multiprocess,
                          identification_progress_callback=identification_progress_callback,
                          fit_progress_callback=fit_progress_callback, return_info=True)
    fitting_method_3d = "gausslq" if fitting_method in ["gausslq", "gausslq-gpu"] else "gaussmle"
    return zfit.zfit(locs, info, calibration=calibration_3d, fitting_method=fitting_method_3d,
                     filter=0, multiprocess=multiprocess, progress_callback=fit_z_progress_callback)
