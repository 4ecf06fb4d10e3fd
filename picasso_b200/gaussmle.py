"""Maximum-likelihood 2-D Gaussian spot fitting on B200.

Drop-in for ``picasso.gaussmle`` (reference picasso/gaussmle.py): same public
functions, signatures, return types and error behaviour; the per-spot numba
kernels (``_mlefit_sigmaxy`` :745, ``_mlefit_sigma`` :533 and their CRLB passes)
are replaced by the CUDA kernel in csrc/mle_fit.cu through ``pb_mle_fit``.
"""
from __future__ import annotations

import threading
from typing import Callable, Literal

import numpy as np
import pandas as pd

from . import _lib

_METHODS = {"sigma": 0, "sigmaxy": 1}


def _method_id(method) -> int:
    if method not in _METHODS:
        raise ValueError("Method not available.")  # reference gaussmle.py:465,513
    return _METHODS[method]


def _as_spots(spots) -> np.ndarray:
    spots = np.ascontiguousarray(spots, dtype=np.float32)
    if spots.ndim != 3 or spots.shape[1] != spots.shape[2]:
        raise ValueError("spots must have shape (N, size, size)")
    return spots


def _fit_into(spots, eps, max_it, method_id, thetas, CRLBs, likelihoods, iterations,
              status=None, progress=None):
    lib = _lib.load()
    _lib.require_gpu()
    n, box, _ = spots.shape
    _lib.check(
        lib.pb_mle_fit(
            n, box, _lib.ptr(spots), float(eps), int(max_it), method_id,
            _lib.ptr(thetas), _lib.ptr(CRLBs), _lib.ptr(likelihoods), _lib.ptr(iterations),
            _lib.ptr(status) if status is not None else None,
            _lib.ptr(progress) if progress is not None else None,
        )
    )


def gaussmle(
    spots,
    eps: float,
    max_it: int,
    method: Literal["sigma", "sigmaxy"] = "sigmaxy",
    progress_callback: Callable[[int], None] | Literal["console"] | None = None,
):
    """Fit Gaussians by MLE to ``spots`` (N, size, size).

    Same contract as reference ``gaussmle.gaussmle`` (gaussmle.py:409-475):
    returns ``(thetas f32 (N,6), CRLBs f32 (N,6), likelihoods f32 (N,),
    iterations i32 (N,))``; ``thetas`` columns are x, y, photons, bg, sx, sy;
    unknown ``method`` raises ``ValueError("Method not available.")``;
    a callable ``progress_callback`` is called with 0..N-1 in order.
    """
    method_id = _method_id(method)
    spots = _as_spots(spots)
    N = len(spots)
    # every entry is written by the kernel (the reference pre-fills CRLBs with inf, :457).  The
    # returned arrays are backed by pooled page-locked memory (_lib.pinned_empty): the device ->
    # host copy is a direct DMA, no staging copy and no first-touch page faults.
    thetas = _lib.pinned_empty((N, 6), np.float32)
    CRLBs = _lib.pinned_empty((N, 6), np.float32)
    likelihoods = _lib.pinned_empty((N,), np.float32)
    iterations = _lib.pinned_empty((N,), np.int32)
    if N:
        _fit_into(spots, eps, max_it, method_id, thetas, CRLBs, likelihoods, iterations)
    if progress_callback == "console":
        from tqdm import tqdm

        with tqdm(total=N, desc="Fitting...", unit="spot") as bar:
            bar.update(N)
    elif callable(progress_callback):
        for i in range(N):
            progress_callback(i)
    return thetas, CRLBs, likelihoods, iterations


def gaussmle_async(
    spots,
    eps: float,
    max_it: int,
    method: Literal["sigma", "sigmaxy"] = "sigmaxy",
):
    """Asynchronous variant (reference gaussmle.py:478-530): returns at once
    with ``(current, thetas, CRLBs, likelihoods, iterations)``; a host thread
    streams the spots through the GPU, fills the shared arrays in place and
    advances ``current[0]`` until it equals N.
    """
    method_id = _method_id(method)
    spots = _as_spots(spots)
    N = len(spots)
    thetas = np.zeros((N, 6), dtype=np.float32)
    CRLBs = np.inf * np.ones((N, 6), dtype=np.float32)
    likelihoods = np.zeros(N, dtype=np.float32)
    iterations = np.zeros(N, dtype=np.int32)
    current = [0]
    _lib.load()
    _lib.require_gpu()

    def _run():
        step = 1 << 20
        for first in range(0, N, step):
            last = min(N, first + step)
            _fit_into(spots[first:last], eps, max_it, method_id, thetas[first:last],
                      CRLBs[first:last], likelihoods[first:last], iterations[first:last])
            current[0] = last

    if N:
        threading.Thread(target=_lib.on_callers_device(_run), name="picasso_b200-mle", daemon=True).start()
    return current, thetas, CRLBs, likelihoods, iterations


def locs_from_fits(
    identifications: pd.DataFrame,
    theta: np.ndarray,
    CRLBs: np.ndarray,
    log_likelihoods: np.ndarray,
    iterations: np.ndarray,
    box: int,
) -> pd.DataFrame:
    """Build the localization table from MLE fit results.

    Same columns, dtypes and ordering as reference ``gaussmle.locs_from_fits``
    (gaussmle.py:957-1037): spot-frame coordinates are shifted by the
    identification pixel minus ``box // 2``; ``lpx/lpy`` are sqrt(CRLB);
    sorted by ``n_id`` when present, else by ``frame``.
    """
    half = int(box / 2)
    f32 = np.float32
    with np.errstate(invalid="ignore"):
        unc = np.sqrt(CRLBs)
        s_big = np.maximum(theta[:, 4], theta[:, 5])
        s_small = np.minimum(theta[:, 4], theta[:, 5])
        ellipticity = (s_big - s_small) / s_big
    columns = {
        "frame": identifications["frame"].to_numpy(dtype=np.uint32),
        "x": (theta[:, 0] + identifications["x"] - half).astype(f32),
        "y": (theta[:, 1] + identifications["y"] - half).astype(f32),
        "photons": theta[:, 2].astype(f32),
        "sx": theta[:, 4].astype(f32),
        "sy": theta[:, 5].astype(f32),
        "bg": theta[:, 3].astype(f32),
        "lpx": unc[:, 0].astype(f32),
        "lpy": unc[:, 1].astype(f32),
        "ellipticity": ellipticity.astype(f32),
        "net_gradient": identifications["net_gradient"].astype(f32),
        "log_likelihood": log_likelihoods.astype(f32),
        "iterations": iterations.astype(np.uint32),
        "photons_unc": unc[:, 2].astype(f32),
        "bg_unc": unc[:, 3].astype(f32),
        "sx_unc": unc[:, 4].astype(f32),
        "sy_unc": unc[:, 5].astype(f32),
    }
    locs = pd.DataFrame(columns)
    if "n_id" in identifications.columns:
        locs["n_id"] = identifications.n_id.astype(np.uint32)
        locs.sort_values(by=["n_id"], kind="quicksort", inplace=True)
    else:
        locs.sort_values(by=["frame"], kind="quicksort", inplace=True)
    return locs


def sigma_uncertainty(sigma, sigma_orth, photons, bg):
    """Standard error of the fitted sigma for the MLE Gaussian/Poisson model (Rieger &
    Stallinga 2014 approximation; reference gaussmle.py:1040-1074)."""
    sa2 = sigma ** 2 + 1 / 12
    tau = (2 * np.pi * sa2 * bg) / photons
    var = (sigma ** 2 / (4 * photons)) * (1 + 8 * tau + np.sqrt((8 * tau) / (1 + 2 * tau)))
    return np.sqrt(var)
