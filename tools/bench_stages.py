"""Stage measurements for the non-headline BASELINE.json configs (3: identify + get_spots +
gausslq end to end; 4: render 50 M localisations; 5: RCC cross-correlation), one JSON line
each, with the HBM roofline of the dominant kernel.  Run on the GPU box:

    python tools/bench_stages.py [identify] [render] [rcc] [--small]

Numbers are kept under profiles/.  (bench.py remains the headline MLE metric.)
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def ev_time(torch, fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gen_movie_device(torch, F, Y, X, per_frame=60, seed=1, dev="cuda"):
    """Config 3 movie on the device: 100 + Poisson(20 + emitters), uint16."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    movie = torch.empty((F, Y, X), dtype=torch.int16, device=dev)
    r = 5
    oy, ox = torch.meshgrid(torch.arange(-r, r + 1, device=dev), torch.arange(-r, r + 1, device=dev),
                            indexing="ij")
    step = 100
    for f0 in range(0, F, step):
        nf = min(step, F - f0)
        mu = torch.full((nf, Y, X), 20.0, device=dev)
        n = nf * per_frame
        fx = 8 + (X - 16) * torch.rand(n, generator=g, device=dev)
        fy = 8 + (Y - 16) * torch.rand(n, generator=g, device=dev)
        ff = torch.arange(nf, device=dev).repeat_interleave(per_frame)
        cx, cy = fx.floor().long(), fy.floor().long()
        px = cx[:, None, None] + ox[None]
        py = cy[:, None, None] + oy[None]
        val = 2000.0 * torch.exp(-0.5 * (((px - fx[:, None, None]) / 1.1) ** 2 +
                                         ((py - fy[:, None, None]) / 1.1) ** 2))
        idx = (ff[:, None, None] * Y + py) * X + px
        mu.view(-1).index_add_(0, idx.reshape(-1), val.reshape(-1))
        movie[f0:f0 + nf] = (100 + torch.poisson(mu, generator=g)).clamp_(0, 65535).to(torch.int32).to(torch.int16)
    return movie.view(torch.uint16)


def stage_identify(torch, small):
    from picasso_b200 import _lib, gausslq, localize

    lib = _lib.load()
    localize._declare(lib)
    vp = C.c_void_p
    lib.pb_identify_dev.argtypes = [vp, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_longlong, C.c_int,
                                    C.c_double, vp, vp, vp, vp, vp, C.c_size_t, vp, vp]
    lib.pb_identify_dev.restype = C.c_int
    F, Y, X = (200, 512, 512) if small else (2000, 512, 512)
    movie = gen_movie_device(torch, F, Y, X)
    cap = F * 256
    fr = torch.empty(cap, dtype=torch.int64, device="cuda"); xs = torch.empty_like(fr); ys = torch.empty_like(fr)
    ng = torch.empty(cap, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        cnt.zero_()
        rc = lib.pb_identify_dev(movie.data_ptr(), 0, F, Y, X, 0, 7, 5000.0, None, fr.data_ptr(),
                                 xs.data_ptr(), ys.data_ptr(), ng.data_ptr(), cap, cnt.data_ptr(), st)
        assert rc == 0
    ms = ev_time(torch, run)
    n_found = int(cnt.item())
    bytes_alg = F * Y * X * 2 + n_found * 28
    ach = bytes_alg / (ms * 1e-3) / 1e9
    if KERNEL_ONLY:
        print(json.dumps({"stage": "identify kernel", "movie": [F, Y, X], "ms": ms,
                          "achieved_GBs": ach, "frac": ach / peaks()}), flush=True)
        return
    # end to end through the Python API with a host movie (H2D inside), incl. get_spots + LQ
    hmovie = movie.cpu().numpy()
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    # warm-up (CUDA context, local-memory pool for the LQ kernel, pandas)
    localize.localize(hmovie[:8], dict(cam), {"Min. Net Gradient": 5000, "Box Size": 7},
                      return_info=False, fitting_method="gausslq")
    # first calls pay lazy module loading and the cached pinned / device buffers: time the second
    first = {}
    for rep in range(2):
        t0 = time.perf_counter()
        ids = localize.identify(hmovie, 5000, 7, return_info=False)
        t1 = time.perf_counter()
        spots = localize.get_spots(hmovie, ids, 7, cam)
        t2 = time.perf_counter()
        theta = gausslq.fit_spots(spots)
        t3 = time.perf_counter()
        if rep == 0:
            first = {"identify_host_movie": t1 - t0, "get_spots": t2 - t1, "gausslq": t3 - t2}
    locs = localize.localize(hmovie, dict(cam), {"Min. Net Gradient": 5000, "Box Size": 7},
                             return_info=False, fitting_method="gausslq")
    t3b = time.perf_counter()
    assert len(locs) == len(ids)
    # CPU oracle on a bounded sample
    import oracle
    oracle.build()
    nf_cpu = 16
    t4 = time.perf_counter()
    ofr, ox, oy, ong = oracle.identify_movie(hmovie[:nf_cpu], 5000, 7)
    t5 = time.perf_counter()
    ns = min(len(spots), 20000)
    oracle.fit_spots_lq(spots[:ns], nthreads=os.cpu_count())
    t6 = time.perf_counter()
    sel = ids["frame"].to_numpy() < nf_cpu
    same = (np.array_equal(ids["x"].to_numpy()[sel], ox) and np.array_equal(ids["y"].to_numpy()[sel], oy))
    print(json.dumps({
        "stage": "identify+get_spots+gausslq (config 3)", "movie": [F, Y, X], "n_identified": n_found,
        "identify_kernel_ms": ms, "identify_fps": F / (ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks(), "unit": "GB/s",
                     "frac": ach / peaks(), "algorithmic_bytes": bytes_alg},
        "e2e_seconds": {"identify_host_movie": t1 - t0, "get_spots": t2 - t1, "gausslq": t3 - t2,
                        "total": t3 - t0, "localize_fused": t3b - t3, "first_call": first},
        "e2e_fps": F / (t3 - t0), "e2e_fps_fused_localize": F / (t3b - t3), "lq_fits_per_s_e2e": len(spots) / (t3 - t2),
        "cpu_oracle": {"identify_fps_1thread": nf_cpu / (t5 - t4),
                       "lq_fits_per_s_allcores": ns / (t6 - t5), "cores": os.cpu_count()},
        "identifications_match_oracle_sample": bool(same)}), flush=True)
    # LQ kernel alone, device resident
    lib.pb_lq_fit_dev.argtypes = [C.c_size_t, C.c_int, vp, vp, vp, vp, vp]
    dsp = torch.from_numpy(spots).cuda()
    dth = torch.empty((len(spots), 6), device="cuda")
    ms_lq = ev_time(torch, lambda: lib.pb_lq_fit_dev(len(spots), 7, dsp.data_ptr(), dth.data_ptr(),
                                                    None, None, st))
    print(json.dumps({"stage": "gausslq kernel (device resident)", "n": len(spots), "ms": ms_lq,
                      "fits_per_s": len(spots) / (ms_lq * 1e-3)}), flush=True)


def stage_localize(torch, small):
    """Config 3 end to end through the fused movie -> localization-table path (pb_localize and
    localize.localize): host movie in (pageable / pinned), locs DataFrame out."""
    from picasso_b200 import _lib, localize

    lib = _lib.load()
    localize._declare(lib)
    F, Y, X = (200, 512, 512) if small else (2000, 512, 512)
    movie = gen_movie_device(torch, F, Y, X)
    hmovie = movie.cpu().numpy()
    del movie
    pin = _lib.PinnedArray(hmovie.shape, hmovie.dtype)
    pin.array[...] = hmovie
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    par = {"Min. Net Gradient": 5000, "Box Size": 7}
    out = {"stage": "localize fused (config 3)", "movie": [F, Y, X], "movie_bytes": int(hmovie.nbytes)}
    for fm in ("gausslq", "gaussmle"):
        localize.localize(hmovie[:8], dict(cam), par, return_info=False, fitting_method=fm)
        for name, mv in (("pageable", hmovie), ("pinned", pin.array)):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                locs = localize.localize(mv, dict(cam), par, return_info=False, fitting_method=fm)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out[f"{fm}_{name}"] = {"seconds": best, "fps": F / best, "n_locs": len(locs),
                                   "movie_GBs": hmovie.nbytes / best / 1e9}
        # the C call alone (no pandas)
        fit = 2 if fm == "gausslq" else 1
        ncol = lib.pb_locs_columns(fit)
        cap = 256 * F
        cols = np.empty((ncol, cap), np.float32)
        found = C.c_size_t(0)
        for name, mv in (("pageable", hmovie), ("pinned", pin.array)):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                rc = lib.pb_localize(_lib.ptr(mv), 0, F, Y, X, 0, 7, 5000.0, None, 100.0, 1.0, 1.0, fit, 1e-3,
                                     100, 0, _lib.ptr(cols), cap, C.byref(found))
                dt = time.perf_counter() - t0
                assert rc == 0
                best = dt if best is None else min(best, dt)
            out[f"{fm}_{name}_c_call"] = {"seconds": best, "fps": F / best, "n_locs": int(found.value),
                                          "movie_GBs": hmovie.nbytes / best / 1e9}
    # two-call path for comparison (ROIs via host)
    t0 = time.perf_counter()
    ids = localize.identify(hmovie, 5000, 7, return_info=False)
    locs2, _ = localize.fit2D(hmovie, [], dict(cam), ids, 7, fitting_method="gausslq")
    out["two_call_identify_fit2D_gausslq_seconds"] = time.perf_counter() - t0
    pin.free()
    print(json.dumps(out), flush=True)


def stage_zfit(torch, small):
    """Astigmatic z fit: pb_zfit_dev on device-resident columns, the Python
    API end to end, and the CPU oracle (bit-identical to scipy's minimize_scalar loop)."""
    from picasso_b200 import _lib, testing, zfit

    lib = _lib.load()
    zfit._declare(lib)
    vp = C.c_void_p
    lib.pb_zfit_dev.argtypes = [C.c_size_t, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int,
                                vp, vp, vp, vp, vp]
    lib.pb_zfit_dev.restype = C.c_int
    n = 1_000_000 if small else 10_000_000
    locs, info, calib = testing.synthetic_zfit_locs(n, 7)
    cx = np.array(calib["X Coefficients"]); cy = np.array(calib["Y Coefficients"])
    d = {k: torch.from_numpy(locs[k].to_numpy()).cuda() for k in ("sx", "sy", "photons", "bg")}
    z = torch.empty(n, device="cuda"); dz = torch.empty(n, device="cuda"); lpz = torch.empty(n, device="cuda")
    nf = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.pb_zfit_dev(n, d["sx"].data_ptr(), d["sy"].data_ptr(), d["photons"].data_ptr(),
                             d["bg"].data_ptr(), None, None, cx.ctypes.data, cy.ctypes.data, 0.79, 130.0, 0,
                             z.data_ptr(), dz.data_ptr(), lpz.data_ptr(), nf.data_ptr(), st)
        assert rc == 0
    ms = ev_time(torch, run)
    alg = n * (16 + 16)          # 4 input + 4 output float32 columns
    t0 = time.perf_counter()
    res, _ = zfit.zfit(locs, info, calibration=dict(calib), fitting_method="gausslq", filter=2)
    t_api = time.perf_counter() - t0
    import oracle
    oracle.build()
    m = 200_000
    t0 = time.perf_counter()
    oz, osq, _, _, onf = oracle.zfit_minimise(locs["sx"].to_numpy()[:m], locs["sy"].to_numpy()[:m], cx, cy)
    t_cpu = time.perf_counter() - t0
    same = bool((z[:m].cpu().numpy() == oz * np.float32(0.79)).all())
    print(json.dumps({"stage": "zfit (8f rank 3)", "n_locs": n, "kernel_ms": ms, "fits_per_s": n / (ms * 1e-3),
                      "mean_nfev": float(nf.float().mean().item()),
                      "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s",
                                   "frac": alg / (ms * 1e-3) / 1e9 / peaks(),
                                   "note": "32 B/localization; the fit is FP64-latency bound (a dependent chain of ~11 target evaluations)"},
                      "python_api_seconds": t_api, "python_api_fits_per_s": n / t_api, "n_kept": len(res),
                      "cpu_oracle": {"n": m, "fits_per_s_1thread": m / t_cpu},
                      "z_bit_identical_to_oracle_sample": same}), flush=True)


def stage_aim(torch, small):
    """AIM drift correction end to end through the Python API, with the CPU
    oracle (same algorithm as the reference: one sort of the concatenated coordinates per shift and
    segment) timed on a reduced instance."""
    from picasso_b200 import aim, testing

    n_frames, side, lpf = (4000, 256, 50.0) if small else (20000, 512, 250.0)
    locs, info, truth = testing.synthetic_aim_locs(n_frames=n_frames, Y=side, X=side, n_clusters=4000,
                                                   locs_per_frame=lpf, seed=12)
    aim.aim(locs[locs["frame"] < 600], [{**info[0], "Frames": 600}], segmentation=100)      # warm-up
    t0 = time.perf_counter()
    und, _, drift = aim.aim(locs, info, segmentation=100)
    t_gpu = time.perf_counter() - t0
    d = drift["x"].to_numpy(); t = truth[:, 0]
    err = float(np.abs((d - d.mean()) - (t - t.mean())).max())
    from oracle import aim_oracle
    nf_cpu = 1000
    sub = locs[locs["frame"] < nf_cpu]
    t0 = time.perf_counter()
    ound, odrift = aim_oracle.aim(sub, [{**info[0], "Frames": nf_cpu}], 100)
    t_cpu = time.perf_counter() - t0
    gsub, _, gdrift = aim.aim(sub, [{**info[0], "Frames": nf_cpu}], segmentation=100)
    same = gdrift["x"].to_numpy().tobytes() == odrift["x"].to_numpy().tobytes()
    print(json.dumps({"stage": "aim (8f rank 4)", "n_locs": len(locs), "frames": n_frames, "image": [side, side],
                      "n_segments": n_frames // 100, "seconds": t_gpu, "locs_per_s": len(locs) / t_gpu,
                      "max_abs_drift_error_px_vs_injected": err,
                      "cpu_oracle": {"n_locs": len(sub), "frames": nf_cpu, "seconds": t_cpu,
                                     "locs_per_s_1thread": len(sub) / t_cpu},
                      "drift_bit_identical_to_oracle_on_sample": bool(same)}), flush=True)


def stage_link(torch, small):
    """postprocess.link: blinking binding sites, GPU link groups + combined
    table through the Python API; the C oracle (the reference's sequential greedy) beside it."""
    from picasso_b200 import postprocess, testing

    nf, sites, side = (2000, 2000, 256) if small else (10000, 5000, 512)
    locs, info = testing.synthetic_link_locs_fast(nf, sites, side, seed=3)
    postprocess.link(locs.iloc[:20000], info)
    t0 = time.perf_counter()
    linked = postprocess.link(locs, info)
    t_gpu = time.perf_counter() - t0
    sl = locs.sort_values(kind="quicksort", by="frame")
    fr, x, y = sl["frame"].to_numpy(), sl["x"].to_numpy(), sl["y"].to_numpy()
    grp = np.zeros(len(sl), np.int32)
    t0 = time.perf_counter()
    lg = postprocess.get_link_groups(fr, x, y, 0.05, 3, grp)
    t_lg = time.perf_counter() - t0
    import oracle
    oracle.build()
    t0 = time.perf_counter()
    olg = oracle.get_link_groups(fr, x, y, 0.05, 3, grp)
    t_cpu = time.perf_counter() - t0
    print(json.dumps({"stage": "link (8f rank 4)", "n_locs": len(locs), "frames": nf, "locs_per_frame": len(locs) / nf,
                      "n_events": len(linked), "link_seconds_python_api": t_gpu, "locs_per_s": len(locs) / t_gpu,
                      "get_link_groups_seconds": t_lg, "cpu_oracle_get_link_groups_seconds_1thread": t_cpu,
                      "link_groups_bit_identical_to_oracle": bool(np.array_equal(lg, olg))}), flush=True)


def stage_render(torch, small):
    from picasso_b200 import _lib, render as pbrender

    lib = _lib.load()
    pbrender._declare(lib)
    n = 5_000_000 if small else 50_000_000
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    x = 512 * torch.rand(n, generator=g, device="cuda")
    y = 512 * torch.rand(n, generator=g, device="cuda")
    lpx = 0.02 + 0.06 * torch.rand(n, generator=g, device="cuda")
    lpy = 0.02 + 0.06 * torch.rand(n, generator=g, device="cuda")
    npx = 10240
    img = torch.empty((npx, npx), device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    wsb = lib.pb_render_workspace_bytes(n, npx, npx)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = {"stage": "render (config 4)", "n_locs": n, "image": [npx, npx]}
    for name, mode, use_ws in (("hist", 0, True), ("gaussian_tiled", 1, True), ("gaussian_direct", 1, False)):
        def run():
            rc = lib.pb_render_dev(n, x.data_ptr(), y.data_ptr(), lpx.data_ptr(), lpy.data_ptr(), 20.0,
                                   0.0, 0.0, 512.0, 512.0, 0.0, mode, img.data_ptr(), npx, npx,
                                   cnt.data_ptr(), ws.data_ptr() if use_ws else None,
                                   wsb if use_ws else 0, st)
            assert rc == 0
        ms = ev_time(torch, run, warm=1, reps=3)
        alg = n * (16 if mode else 8) + npx * npx * 4
        out[name] = {"ms": ms, "locs_per_s": n / (ms * 1e-3), "achieved_GBs": alg / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peaks(), "algorithmic_bytes": alg}
    if KERNEL_ONLY:
        print(json.dumps(out), flush=True)
        return
    # end to end through the Python API (host arrays in, host image out)
    import pandas as pd
    import warnings
    m = min(n, 10_000_000)
    locs = pd.DataFrame({"x": x[:m].cpu().numpy(), "y": y[:m].cpu().numpy(),
                         "lpx": lpx[:m].cpu().numpy(), "lpy": lpy[:m].cpu().numpy()})
    info = [{"Height": 512, "Width": 512, "Frames": 1, "Pixelsize": 130}]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pbrender.render(locs.iloc[:1000], info, oversampling=20, blur_method="gaussian")
        t0 = time.perf_counter()
        k, image = pbrender.render(locs, info, oversampling=20, blur_method="gaussian")
        t1 = time.perf_counter()
    out["e2e_python_api"] = {"n_locs": m, "seconds": t1 - t0, "locs_per_s": m / (t1 - t0)}
    import oracle
    oracle.build()
    mc = 1_000_000
    sub = {k_: v.to_numpy()[:mc] for k_, v in locs.items()}
    t0 = time.perf_counter()
    kc, oimg = oracle.render(sub, info, oversampling=20, blur_method="gaussian")
    t1 = time.perf_counter()
    out["cpu_oracle"] = {"n_locs": mc, "locs_per_s_1thread": mc / (t1 - t0)}
    # parity of the sub-sample
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kg, gimg = pbrender.render(locs.iloc[:mc], info, oversampling=20, blur_method="gaussian")
    big = oimg > 1e-3 * oimg.max()
    out["parity_vs_oracle_1M"] = {"n_equal": kg == kc,
                                  "max_rel_on_bright_pixels": float(np.max(np.abs(gimg[big] - oimg[big]) / oimg[big]))}
    print(json.dumps(out), flush=True)


def stage_rcc(torch, small):
    from picasso_b200 import _lib, imageprocess

    lib = _lib.load()
    imageprocess._declare(lib)
    vp = C.c_void_p
    lib.pb_rcc_spectra_dev.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.pb_rcc_windows_dev.argtypes = [C.c_int, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                                       C.c_int, vp, C.c_int, vp, C.c_size_t, vp]
    n_seg, Y, X = (40, 2048, 2048) if small else (200, 4096, 4096)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    seg = torch.zeros((n_seg, Y, X), device="cuda")
    # sparse blobs shifted a little per segment
    npts = 20000
    py = torch.randint(16, Y - 16, (npts,), generator=g, device="cuda")
    px = torch.randint(16, X - 16, (npts,), generator=g, device="cuda")
    for s in range(n_seg):
        sy, sx = int(round(3 * math.sin(s / 7))), int(round(s / 20))
        seg[s].index_put_((py + sy, px + sx), torch.ones(npts, device="cuda"), accumulate=True)
    spec = torch.empty((n_seg, Y, X // 2 + 1, 2), device="cuda")
    sums = torch.empty(n_seg, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    t_spec = ev_time(torch, lambda: lib.pb_rcc_spectra_dev(n_seg, Y, X, seg.data_ptr(), spec.data_ptr(),
                                                          sums.data_ptr(), st), warm=1, reps=2)
    pi, pj = np.triu_indices(n_seg, 1)
    n_pairs = len(pi)
    dpi = torch.from_numpy(pi.astype(np.int32)).cuda(); dpj = torch.from_numpy(pj.astype(np.int32)).cuda()
    batch = 8
    wsb = batch * (Y * (X // 2 + 1) * 8 + Y * X * 4)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    H = W = 32
    win = torch.empty((n_pairs, H, W), device="cuda")
    Y0, X0 = (Y - 32) // 2, (X - 32) // 2
    lib.pb_rcc_set_mode.argtypes = [C.c_int]
    # CPU: numpy xcorr for one pair (f64, pocketfft)
    a = seg[0].cpu().numpy().astype(np.float64); b = seg[1].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter()
    xc = np.fft.fftshift(np.real(np.fft.ifft2(np.fft.fft2(a) * np.conj(np.fft.fft2(b))))) / np.sqrt(a.size)
    t_cpu = time.perf_counter() - t0
    ref = xc[Y0:Y0 + 32, X0:X0 + 32]
    spec_b = Y * (X // 2 + 1) * 8
    res = {}
    ref_win = None
    for mode, name, fft in ((1, "pruned_fft", "1"), (1, "pruned_direct", "0"), (0, "cufft", "1")):
        os.environ["PB_RCC_FFT"] = fft
        assert lib.pb_rcc_set_mode(mode) == 0
        npr = n_pairs if (name == "pruned_fft" or small) else min(n_pairs, 4000)   # slower paths: timed sample
        win.zero_()
        lib.pb_rcc_windows_dev(min(npr, 64), dpi.data_ptr(), dpj.data_ptr(), Y, X, spec.data_ptr(), Y0, X0,
                               H, W, win.data_ptr(), batch, ws.data_ptr(), wsb, st)      # warm-up
        torch.cuda.synchronize()
        reps = []
        for _ in range(3 if name == "pruned_fft" else 1):
            t0 = time.perf_counter()
            rc = lib.pb_rcc_windows_dev(npr, dpi.data_ptr(), dpj.data_ptr(), Y, X, spec.data_ptr(), Y0, X0, H, W,
                                        win.data_ptr(), batch, ws.data_ptr(), wsb, st)
            torch.cuda.synchronize()
            reps.append(time.perf_counter() - t0)
            assert rc == 0
        t_pairs = min(reps)
        if mode == 1:
            # each pair reads both half-spectra once and writes H x (X/2+1) coefficients
            traffic = npr * (2 * spec_b + 2 * H * (X // 2 + 1) * 8)
            flops = npr * Y * (X // 2 + 1) * (8 * H + 6) if fft == "0" else None
            note = "algorithmic: both half-spectra read once per pair (mostly from L2 with pair tiling)"
        else:
            traffic = npr * (3 * spec_b + 2 * Y * X * 4)   # mul: 2 reads + 1 write; C2R: >= 1 read + 1 write
            flops = None
            note = "minimum traffic of multiply + C2R passes; cuFFT does more than one pass"
        gwin = win[0].cpu().numpy()
        cur = win[:min(npr, 4000)].clone()
        if ref_win is None:
            ref_win = cur
            dev_vs_first = 0.0
        else:
            m = min(len(cur), len(ref_win))
            dev_vs_first = float((cur[:m] - ref_win[:m]).abs().max())
        res[name] = {"pairs_timed": int(npr), "seconds": t_pairs, "pairs_per_s": npr / t_pairs,
                     "all_pairs_seconds": t_pairs * n_pairs / npr,
                     "roofline": {"bound": "hbm", "achieved": traffic / t_pairs / 1e9, "peak": peaks(),
                                  "unit": "GB/s", "frac": traffic / t_pairs / 1e9 / peaks(), "note": note},
                     "fp32_tflops": None if flops is None else flops / t_pairs / 1e12,
                     "window_max_abs_err_vs_numpy_f64": float(np.abs(gwin - ref).max()),
                     "max_abs_dev_vs_pruned_fft_windows": dev_vs_first, "repeat_seconds": reps}
    os.environ.pop("PB_RCC_FFT", None)
    lib.pb_rcc_set_mode(-1)
    print(json.dumps({
        "stage": "rcc (config 5)", "n_seg": n_seg, "image": [Y, X], "n_pairs": n_pairs, "window": [H, W],
        "forward_r2c_all_segments_ms": t_spec, **res,
        "cpu_numpy_xcorr_one_pair_s": t_cpu, "cpu_all_pairs_extrapolated_h": t_cpu * n_pairs / 3600,
        "window_peak": float(ref.max())}), flush=True)


def stage_undrift(torch, small):
    """End-to-end postprocess.undrift through the Python API (host locs DataFrame in, drift
    out): device-rendered segments + cuFFT RCC + batched host peak fits + spline."""
    import warnings

    from picasso_b200 import postprocess, testing

    n_frames, Y, X, seg = (4000, 1024, 1024, 100) if small else (12000, 2048, 2048, 100)
    locs, info, truth = testing.synthetic_drift_locs(n_frames, Y, X, n_clusters=3000,
                                                     locs_per_frame=500, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        postprocess.undrift(locs.iloc[:20000], [{**info[0], "Frames": 400}], 100, display=False,
                            segmentation_callback=lambda i: None, rcc_callback=lambda i: None)
        t0 = time.perf_counter()
        drift, und = postprocess.undrift(locs, info, seg, display=False,
                                         segmentation_callback=lambda i: None,
                                         rcc_callback=lambda i: None)
        t1 = time.perf_counter()
    n_seg = postprocess.n_segments(info, seg)
    err = []
    for k, col in enumerate(("x", "y")):
        est = drift[col].to_numpy() - drift[col].mean()
        tru = truth[:, k] - truth[:, k].mean()
        err.append(float(np.abs(est - tru).max()))
    print(json.dumps({"stage": "postprocess.undrift end to end (Python API)", "frames": n_frames,
                      "image": [Y, X], "n_locs": len(locs), "n_segments": n_seg,
                      "n_pairs": n_seg * (n_seg - 1) // 2, "seconds": t1 - t0,
                      "max_abs_drift_error_px_vs_injected": err}), flush=True)


KERNEL_ONLY = "--kernel-only" in sys.argv

if __name__ == "__main__":
    import torch

    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    small = "--small" in sys.argv
    which = args or ["identify", "render", "rcc"]
    for w in which:
        {"identify": stage_identify, "localize": stage_localize, "zfit": stage_zfit, "aim": stage_aim, "link": stage_link, "render": stage_render, "rcc": stage_rcc,
         "undrift": stage_undrift}[w](torch, small)
