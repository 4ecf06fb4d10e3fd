"""Deterministic synthetic inputs for the BASELINE.json configurations
(SURVEY.md section 8d).  Used by tests/ and bench.py; not part of the hot path.
"""
from __future__ import annotations

import math

import numpy as np


def _erf(x):
    # vectorised erf without scipy dependency at import time
    from scipy.special import erf

    return erf(x)


def synthetic_spots(n: int, box: int = 7, seed=0, return_truth: bool = False):
    """Config 1/2 spots: integrated-pixel Gaussian + Poisson noise, float32.

    x0, y0 ~ U(c-0.5, c+0.5), sigma_x, sigma_y ~ U(0.9, 1.3), photons ~
    U(500, 5000), bg ~ U(5, 30); c = box // 2.
    """
    rng = np.random.default_rng(seed)
    c = box // 2
    x0 = rng.uniform(c - 0.5, c + 0.5, n)
    y0 = rng.uniform(c - 0.5, c + 0.5, n)
    sx = rng.uniform(0.9, 1.3, n)
    sy = rng.uniform(0.9, 1.3, n)
    ph = rng.uniform(500, 5000, n)
    bg = rng.uniform(5, 30, n)
    i = np.arange(box)[None, :]

    def dE(mu, s):
        a = (i - mu[:, None] + 0.5) / (math.sqrt(2) * s[:, None])
        b = (i - mu[:, None] - 0.5) / (math.sqrt(2) * s[:, None])
        return 0.5 * (_erf(a) - _erf(b))

    ex, ey = dE(x0, sx), dE(y0, sy)
    mu = ph[:, None, None] * ey[:, :, None] * ex[:, None, :] + bg[:, None, None]
    spots = rng.poisson(mu).astype(np.float32)
    if return_truth:
        return spots, np.stack([x0, y0, ph, bg, sx, sy], 1)
    return spots


def synthetic_spots_chunked(n: int, box: int = 7, chunk: int = 100_000, seed: int = 0, out=None):
    """Config 2: n spots generated in chunks with seeds SeedSequence(seed).spawn(k)."""
    nchunks = (n + chunk - 1) // chunk
    seeds = np.random.SeedSequence(seed).spawn(nchunks)
    if out is None:
        out = np.empty((n, box, box), np.float32)
    for k in range(nchunks):
        lo, hi = k * chunk, min(n, (k + 1) * chunk)
        out[lo:hi] = synthetic_spots(hi - lo, box, seeds[k])
    return out


def synthetic_movie(frames: int, Y: int, X: int, emitters_per_frame: int = 60, seed: int = 1,
                    sigma: float = 1.1, amplitude: float = 2000.0, baseline: int = 100,
                    bg: float = 20.0, margin: int = 8, return_truth: bool = False):
    """Config 3 movie: uint16 frames = baseline + Poisson(bg + emitters);
    emitters are point-sampled Gaussians with peak `amplitude`, width `sigma`, at uniform
    positions at least `margin` px from the border."""
    rng = np.random.default_rng(seed)
    movie = np.empty((frames, Y, X), np.uint16)
    truth = []
    yy = np.arange(Y, dtype=np.float64)[:, None]
    xx = np.arange(X, dtype=np.float64)[None, :]
    r = int(math.ceil(5 * sigma))
    for f in range(frames):
        mu = np.full((Y, X), bg, np.float64)
        ex = rng.uniform(margin, X - margin, emitters_per_frame)
        ey = rng.uniform(margin, Y - margin, emitters_per_frame)
        for x0, y0 in zip(ex, ey):
            y_lo, y_hi = max(0, int(y0) - r), min(Y, int(y0) + r + 1)
            x_lo, x_hi = max(0, int(x0) - r), min(X, int(x0) + r + 1)
            gy = np.exp(-0.5 * ((yy[y_lo:y_hi] - y0) / sigma) ** 2)
            gx = np.exp(-0.5 * ((xx[:, x_lo:x_hi] - x0) / sigma) ** 2)
            mu[y_lo:y_hi, x_lo:x_hi] += amplitude * gy * gx
        movie[f] = np.minimum(baseline + rng.poisson(mu), 65535).astype(np.uint16)
        truth.append(np.stack([np.full_like(ex, f), ex, ey], 1))
    if return_truth:
        return movie, np.concatenate(truth)
    return movie


def synthetic_drift_locs(n_frames: int, Y: int, X: int, n_clusters: int = 40,
                         locs_per_frame: float = 5.0, seed: int = 3, jitter: float = 0.05,
                         lp: float = 0.05, amp_x: float = 1.0, amp_y: float = 0.7):
    """Config 5 style localizations: clusters with an analytic drift
    x += amp_x * sin(2 pi t / (n_frames / 2)), y += amp_y * (t / n_frames - 0.5)."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    n = int(n_frames * locs_per_frame)
    frame = np.sort(rng.integers(0, n_frames, n)).astype(np.uint32)
    cx = rng.uniform(6, X - 6, n_clusters)
    cy = rng.uniform(6, Y - 6, n_clusters)
    which = rng.integers(0, n_clusters, n)
    t = frame.astype(np.float64)
    dx = amp_x * np.sin(2 * np.pi * t / (n_frames / 2.0))
    dy = amp_y * (t / n_frames - 0.5)
    x = cx[which] + dx + rng.normal(0, jitter, n)
    y = cy[which] + dy + rng.normal(0, jitter, n)
    locs = pd.DataFrame({
        "frame": frame, "x": x.astype(np.float32), "y": y.astype(np.float32),
        "lpx": np.full(n, lp, np.float32), "lpy": np.full(n, lp, np.float32),
    })
    info = [{"Height": Y, "Width": X, "Frames": n_frames, "Pixelsize": 130}]
    drift = np.stack([amp_x * np.sin(2 * np.pi * np.arange(n_frames) / (n_frames / 2.0)),
                      amp_y * (np.arange(n_frames) / n_frames - 0.5)], 1)
    return locs, info, drift


def synthetic_zfit_locs(n: int, seed: int = 11):
    """2-D fitted localizations of an astigmatic 3-D acquisition with a degree-6 calibration
    (coefficient order z^6 .. z^0 like picasso's calibration YAML): returns (locs, info,
    calibration).  A few localizations have widths outside the calibrated range."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    cx = [1.0e-19, -2.0e-16, 1.0e-12, 2.0e-10, 1.5e-6, 1.2e-3, 1.30]
    cy = [-1.5e-19, 1.0e-16, 1.2e-12, -2.5e-10, 1.4e-6, -1.1e-3, 1.32]
    z = rng.uniform(-350, 350, n)
    wx = np.polyval(cx, z)
    wy = np.polyval(cy, z)
    sx = (wx + rng.normal(0, 0.03, n)).astype(np.float32)
    sy = (wy + rng.normal(0, 0.03, n)).astype(np.float32)
    k = max(1, n // 100)
    sx[:k] = rng.uniform(2.5, 3.5, k).astype(np.float32)        # outside the calibration
    sy[k:2 * k] = rng.uniform(0.3, 0.6, k).astype(np.float32)
    photons = rng.uniform(500, 5000, n).astype(np.float32)
    bg = rng.uniform(5, 30, n).astype(np.float32)
    locs = pd.DataFrame({
        "frame": np.sort(rng.integers(0, 1000, n)).astype(np.uint32),
        "x": rng.uniform(1, 63, n).astype(np.float32), "y": rng.uniform(1, 63, n).astype(np.float32),
        "photons": photons, "sx": sx, "sy": sy, "bg": bg,
        "lpx": rng.uniform(0.01, 0.05, n).astype(np.float32),
        "lpy": rng.uniform(0.01, 0.05, n).astype(np.float32),
        "ellipticity": (np.abs(sx - sy) / np.maximum(sx, sy)).astype(np.float32),
        "net_gradient": rng.uniform(5000, 20000, n).astype(np.float32),
        "sx_unc": (sx / np.sqrt(photons) * rng.uniform(0.9, 1.5, n)).astype(np.float32),
        "sy_unc": (sy / np.sqrt(photons) * rng.uniform(0.9, 1.5, n)).astype(np.float32),
    })
    info = [{"Width": 64, "Height": 64, "Frames": 1000, "Pixelsize": 130}]
    calib = {"X Coefficients": cx, "Y Coefficients": cy, "Magnification factor": 0.79}
    return locs, info, calib


def synthetic_aim_locs(n_frames: int = 2000, Y: int = 64, X: int = 64, n_clusters: int = 60,
                       locs_per_frame: float = 10.0, seed: int = 4, with_z: bool = False):
    """Localizations for AIM undrifting: clusters with a slow analytic drift (well inside the
    default 60 nm search radius per 100-frame segment), optional z column in nm."""
    locs, info, drift = synthetic_drift_locs(n_frames, Y, X, n_clusters=n_clusters,
                                             locs_per_frame=locs_per_frame, seed=seed, jitter=0.04,
                                             amp_x=0.45, amp_y=0.6)
    if with_z:
        rng = np.random.default_rng(seed + 1000)
        n = len(locs)
        cz = rng.uniform(-300, 300, n_clusters)
        which = rng.integers(0, n_clusters, n)
        t = locs["frame"].to_numpy().astype(np.float64)
        dz = 40.0 * np.sin(2 * np.pi * t / n_frames)
        locs["z"] = (cz[which] + dz + rng.normal(0, 8.0, n)).astype(np.float32)
    return locs, info, drift


def synthetic_link_locs(n_frames: int = 600, n_sites: int = 150, side: int = 64, seed: int = 5,
                        with_group: bool = False, f64_xy: bool = False):
    """Blinking binding sites (on-times ~6 frames, dark gaps inside an event up to 2 frames, sites
    >= 0.5 px apart) for postprocess.link: frame-sorted DataFrame + info."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    sites = []
    while len(sites) < n_sites:
        p = rng.uniform(3, side - 3, 2)
        if all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 > 0.25 for q in sites):
            sites.append(p)
    rows = []
    for k, (sx_, sy_) in enumerate(sites):
        t = int(rng.integers(0, 40))
        while t < n_frames:
            length = 1 + int(rng.geometric(1 / 6.0))
            for f in range(t, min(t + length, n_frames)):
                if rng.random() < 0.85:                                # missed frames inside an event
                    rows.append((f, sx_ + rng.normal(0, 0.012), sy_ + rng.normal(0, 0.012), k))
            t += length + int(rng.integers(8, 120))
    rows.sort(key=lambda r: r[0])
    a = np.array(rows)
    n = len(a)
    xy = np.float64 if f64_xy else np.float32
    locs = pd.DataFrame({
        "frame": a[:, 0].astype(np.uint32), "x": a[:, 1].astype(xy), "y": a[:, 2].astype(xy),
        "photons": rng.uniform(500, 5000, n).astype(np.float32), "sx": rng.uniform(0.9, 1.3, n).astype(np.float32),
        "sy": rng.uniform(0.9, 1.3, n).astype(np.float32), "bg": rng.uniform(5, 30, n).astype(np.float32),
        "lpx": rng.uniform(0.008, 0.02, n).astype(np.float32), "lpy": rng.uniform(0.008, 0.02, n).astype(np.float32),
        "ellipticity": rng.uniform(0, 0.2, n).astype(np.float32),
        "net_gradient": rng.uniform(5000, 20000, n).astype(np.float32),
    })
    if with_group:
        locs["group"] = (a[:, 3].astype(np.int32) % 7)
    info = [{"Height": side, "Width": side, "Frames": n_frames, "Pixelsize": 130}]
    return locs, info


def synthetic_link_locs_fast(n_frames: int, n_sites: int, side: float, seed: int = 3):
    """Vectorised variant of ``synthetic_link_locs`` for large benchmarks: sites on a jittered grid
    (>= 0.5 px apart), geometric on-times (mean 6 frames), 15 % missed frames."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    k = int(np.ceil(np.sqrt(n_sites)))
    pitch = side / k
    assert pitch >= 0.7, "sites would be closer than 0.5 px"
    gy, gx = np.divmod(np.arange(n_sites), k)
    sx = (gx + 0.5) * pitch + rng.uniform(-0.1, 0.1, n_sites)
    sy = (gy + 0.5) * pitch + rng.uniform(-0.1, 0.1, n_sites)
    n_ev = max(1, n_frames // 50)
    on = rng.geometric(1 / 6.0, (n_sites, n_ev))
    off = rng.integers(8, 90, (n_sites, n_ev))
    start = np.cumsum(on + off, axis=1) - on - rng.integers(0, 8, (n_sites, 1))
    site = np.repeat(np.arange(n_sites), n_ev)
    start, on = start.ravel(), on.ravel()
    reps = np.repeat(np.arange(len(on)), on)
    within = np.arange(on.sum()) - np.repeat(np.cumsum(on) - on, on)
    frame = start[reps] + within
    s = site[reps]
    keep = (frame >= 0) & (frame < n_frames) & (rng.random(len(frame)) < 0.85)
    frame, s = frame[keep], s[keep]
    order = np.argsort(frame, kind="stable")
    frame, s = frame[order], s[order]
    n = len(frame)
    f32 = np.float32
    locs = pd.DataFrame({
        "frame": frame.astype(np.uint32), "x": (sx[s] + rng.normal(0, 0.012, n)).astype(f32),
        "y": (sy[s] + rng.normal(0, 0.012, n)).astype(f32), "photons": rng.uniform(500, 5000, n).astype(f32),
        "sx": rng.uniform(0.9, 1.3, n).astype(f32), "sy": rng.uniform(0.9, 1.3, n).astype(f32),
        "bg": rng.uniform(5, 30, n).astype(f32), "lpx": rng.uniform(0.008, 0.02, n).astype(f32),
        "lpy": rng.uniform(0.008, 0.02, n).astype(f32), "ellipticity": rng.uniform(0, 0.2, n).astype(f32),
        "net_gradient": rng.uniform(5000, 20000, n).astype(f32),
    })
    return locs, [{"Height": int(side), "Width": int(side), "Frames": n_frames, "Pixelsize": 130}]
