"""The few ``picasso.lib`` helpers the hot path needs (reference picasso/lib.py):
``get_from_metadata`` :878-920, ``ensure_sanity`` :1786-1832 and ``minimize_shifts`` :2034-2078.  Host-side numpy;
no GUI code."""
from __future__ import annotations

from typing import Any

import numpy as np


def get_from_metadata(info, key: Any, default=None, *, raise_error: bool = False):
    """Search the metadata (dict, or list of dicts from last to first) for ``key``
    (reference lib.py:878-920; a falsy value counts as missing in the list form)."""
    if isinstance(info, dict):
        if raise_error and key not in info:
            raise KeyError(f"Key '{key}' not found in metadata.")
        return info.get(key, default)
    if isinstance(info, list):
        for entry in reversed(info):
            value = entry.get(key)
            if value:
                return value
        if raise_error:
            raise KeyError(f"Key '{key}' not found in metadata.")
        return default
    raise ValueError("info must be a dict or a list of dicts.")


def minimize_shifts(shifts_x, shifts_y, shifts_z=None):
    """Least-squares chain of pairwise shifts (RCC; reference lib.py:2034-2078): the reference
    solves ``A d = r`` with ``A[pair(i, j), i:j] = 1`` by ``pinv`` of the (n(n-1)/2, n-1) matrix and
    returns the cumulative shifts (leading 0) as ``(shift_y, shift_x[, shift_z])``.

    ``A`` has full column rank, so ``pinv(A) r`` is the solution of the normal equations, and both
    sides have closed forms: ``(A^T A)[k, l] = (min(k, l) + 1) (n - 1 - max(k, l))`` (pairs with
    i <= min, j > max) and ``(A^T r)[k] = sum_{i <= k < j} r_ij`` (two prefix sums of the shift
    matrix).  O(n^2) instead of the SVD of a 19 900 x 199 matrix at 200 segments (0.2-1 s);
    agrees with the pinv formulation to 2e-14 (tests pin it to the reference at 1e-9)."""
    n = shifts_x.shape[0]
    stacks = [shifts_y, shifts_x] + ([shifts_z] if shifts_z is not None else [])
    m = n - 1
    if m < 1:
        return tuple(np.zeros(max(n, 1))[:n] for _ in stacks)
    k = np.arange(m)
    gram = (np.minimum.outer(k, k) + 1.0) * (n - 1 - np.maximum.outer(k, k))
    out = []
    for sh in stacks:
        upper = np.triu(np.asarray(sh, dtype=np.float64), 1)
        below = np.cumsum(upper, axis=0)                          # sum over i <= k
        right = np.cumsum(below[:, ::-1], axis=1)[:, ::-1]        # ... and over j' >= j
        rhs = right[k, k + 1]
        d = np.linalg.solve(gram, rhs)
        out.append(np.insert(np.cumsum(d), 0, 0))
    return tuple(out)


def _minimize_shifts_pinv(shifts_x, shifts_y, shifts_z=None):
    """The reference's literal formulation (lib.py:2034-2078), kept for the equivalence test."""
    n = shifts_x.shape[0]
    pairs = [(i, j) for i in range(n - 1) for j in range(i + 1, n)]
    stacks = [shifts_y, shifts_x] + ([shifts_z] if shifts_z is not None else [])
    rij = np.zeros((len(pairs), len(stacks)))
    A = np.zeros((len(pairs), n - 1))
    for row, (i, j) in enumerate(pairs):
        for col, sh in enumerate(stacks):
            rij[row, col] = sh[i, j]
        A[row, i:j] = 1
    Dj = np.dot(np.linalg.pinv(A), rij)
    return tuple(np.insert(np.cumsum(Dj[:, c]), 0, 0) for c in range(len(stacks)))


def ensure_sanity(locs, info):
    """Drop localizations with inf / NaN entries, outside the image or with negative
    precisions / sizes (reference lib.py:1786-1832); ``info`` must hold Width, Height, Frames.

    The reference replaces inf by NaN, drops NaN rows and then applies eleven boolean filters one
    after the other, copying the table each time; the same rows are selected here with one
    combined mask (on 10 M localizations: 0.25 s instead of 1.1 s)."""
    for key in ["Width", "Height", "Frames"]:
        if get_from_metadata(info, key) is None:
            raise KeyError(f"Metadata is missing required key: '{key}'")
    keep = np.ones(len(locs), dtype=bool)
    for name in locs.columns:
        col = locs[name].to_numpy()
        if col.dtype.kind == "f":
            keep &= np.isfinite(col)
        elif col.dtype.kind not in "iub":
            keep &= ~np.asarray(locs[name].isna())            # object / datetime columns
    with np.errstate(invalid="ignore"):
        keep &= locs["x"].to_numpy() < get_from_metadata(info, "Width")
        keep &= locs["y"].to_numpy() < get_from_metadata(info, "Height")
        for attr in ["x", "y", "lpx", "lpy", "lpz", "photons", "ellipticity", "sx", "sy"]:
            if attr in locs.columns:
                keep &= locs[attr].to_numpy() >= 0
    return locs[keep]
