"""Stage measurements of BASELINE.json configs 3 / 4 / 5 for bench.py's ``stages`` block, at the
world size the bench was launched with (one process per GPU, ``picasso_b200.distributed``):

  localize_config3   2000 x 512 x 512 uint16 movie -> localization table (identify + get_spots +
                     gausslq + locs_from_fits fused on the GPU); frames shard by contiguous block
  render_config4     50 M localizations -> 10240 x 10240 Gaussian render (oversampling 20); the image
                     shards by row bands, one NCCL all-to-all of the bucketed localizations
  undrift_config5    20 M localizations / 20 000 frames / 4096^2 -> 200 segments, 19 900 pair
                     cross-correlations; segments shard for render + R2C, spectra all-gathered,
                     pairs shard by L2 tile

Every stage reports
  seconds       device-resident: inputs in HBM when the clock starts, result in HBM (localize,
                render band) or on the host (undrift shifts + drift: the device part ends in a few KB)
  e2e_seconds   through the host-facing sharded call: this rank's share of the input in host
                memory, H2D and D2H inside the timed region
both as the MAX over ranks of the best of `repeats` synchronous calls (barrier + device
synchronize on both sides; the calls themselves are synchronous, so host timers see all of the
device work), the per-kernel / per-phase split from CUDA events, a roofline entry for the dominant
kernel, and a parity flag: N > 1 against the same stage run on ONE GPU inside the same process
(bit-equal tables / 1e-4 relative pixels / 1e-6 px shifts), N = 1 against the CPU oracle or the
injected ground truth on a bounded sample.  Synthetic data are generated on the device from
per-chunk seeds, so the union over ranks is the same data set for every N.
"""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CAM = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
PARAMS = {"Min. Net Gradient": 5000, "Box Size": 7}


class Ctx:
    def __init__(self, torch, dist, rank, world, dev, peak_gbs, repeats=3):
        self.torch, self.dist, self.rank, self.world, self.dev = torch, dist, rank, world, dev
        self.peak = peak_gbs
        self.repeats = repeats

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, warmup=1):
        """best-of-repeats wall time of a synchronous call, max over ranks"""
        for _ in range(warmup):
            fn()
        best = float("inf")
        for _ in range(self.repeats):
            self.barrier()
            t0 = time.perf_counter()
            fn()
            self.torch.cuda.synchronize(self.dev)
            best = min(best, time.perf_counter() - t0)
        return self.max_over_ranks(best)

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag):
        if self.world == 1:
            return bool(flag)
        t = self.torch.tensor([1.0 if flag else 0.0], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def sum_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def _bounds(n, world):
    return [(n * r) // world for r in range(world + 1)]


# ---------------------------------------------------------------------------------------------
# config 3: fused localize
# ---------------------------------------------------------------------------------------------
def gen_movie_chunk(torch, chunk_index, frames, Y, X, dev, emitters=60, sigma=1.1, amplitude=2000.0,
                    baseline=100, bg=20.0, margin=8):
    """Config-3 frames (same model as picasso_b200.testing.synthetic_movie) generated
    on the device: uint16 = baseline + Poisson(bg + sum of point-sampled Gaussians of peak
    `amplitude`, width `sigma`, at uniform positions >= `margin` px from the border)."""
    g = torch.Generator(device=dev)
    g.manual_seed(31_000 + chunk_index)
    r = int(math.ceil(5 * sigma))
    k = 2 * r + 1
    ex = margin + (X - 2 * margin) * torch.rand((frames, emitters), generator=g, device=dev, dtype=torch.float64)
    ey = margin + (Y - 2 * margin) * torch.rand((frames, emitters), generator=g, device=dev, dtype=torch.float64)
    off = torch.arange(-r, r + 1, device=dev)
    py = ey.long()[..., None] + off                      # (F, E, k) pixel rows
    px = ex.long()[..., None] + off
    gy = torch.exp(-0.5 * ((py.double() - ey[..., None]) / sigma) ** 2)
    gx = torch.exp(-0.5 * ((px.double() - ex[..., None]) / sigma) ** 2)
    val = amplitude * gy[..., :, None] * gx[..., None, :]            # (F, E, k, k)
    ok = ((py >= 0) & (py < Y))[..., :, None] & ((px >= 0) & (px < X))[..., None, :]
    fidx = torch.arange(frames, device=dev)[:, None, None, None]
    lin = (fidx * Y + py.clamp(0, Y - 1)[..., :, None]) * X + px.clamp(0, X - 1)[..., None, :]
    mu = torch.full((frames * Y * X,), bg, dtype=torch.float32, device=dev)
    mu.index_add_(0, lin[ok].reshape(-1), val[ok].float().reshape(-1))
    cnt = torch.poisson(mu, generator=g)
    # int16 storage, read by the library as uint16 (counts stay far below 32768)
    out = torch.clamp(cnt + baseline, max=32767).to(torch.int32).to(torch.int16)
    return out.reshape(frames, Y, X)


def stage_localize(ctx, frames=2000, Y=512, X=512, chunk=100):
    torch, dist = ctx.torch, ctx.dist
    from picasso_b200 import _lib, distributed as pbd, localize as pbl

    lib = pbl._lib_ready()           # declares the ctypes signatures of the localize entry points
    nchunks = frames // chunk
    cb = _bounds(nchunks, ctx.world)
    my_chunks = range(cb[ctx.rank], cb[ctx.rank + 1])
    f0 = cb[ctx.rank] * chunk
    movie = (torch.cat([gen_movie_chunk(torch, c, chunk, Y, X, ctx.dev) for c in my_chunks])
             if len(my_chunks) else torch.empty((0, Y, X), dtype=torch.int16, device=ctx.dev))
    nf = int(movie.shape[0])

    def gather(cols):
        """all-gather of the finished column blocks (padded): every rank ends with the whole table"""
        if ctx.world == 1:
            return cols
        cnt = torch.tensor([cols.shape[1]], dtype=torch.int64, device=ctx.dev)
        counts = torch.empty(ctx.world, dtype=torch.int64, device=ctx.dev)
        dist.all_gather_into_tensor(counts, cnt)
        counts = counts.cpu().tolist()
        nmax = max(max(counts), 1)
        pad = torch.zeros((cols.shape[0], nmax), dtype=torch.float32, device=ctx.dev)
        pad[:, : cols.shape[1]] = cols
        out = torch.empty((ctx.world, cols.shape[0], nmax), dtype=torch.float32, device=ctx.dev)
        dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
        return torch.cat([out[r][:, : counts[r]] for r in range(ctx.world)], dim=1)

    result = {}

    def step_dev():
        cols = (pbd.localize_device(torch, movie, f0, CAM, PARAMS, fitting_method="gausslq") if nf
                else torch.empty((11, 0), dtype=torch.float32, device=ctx.dev))
        result["cols"] = gather(cols)

    seconds = ctx.timed(step_dev)
    table = result["cols"]
    n_locs = int(table.shape[1])

    # ---- end to end: this rank's frame block in pinned host memory -> columns on the host ----
    import ctypes as C

    hm = _lib.PinnedArray((max(nf, 1), Y, X), np.uint16)
    if nf:
        torch.from_numpy(hm.array.view(np.int16)).copy_(movie.cpu())
    ncols = 11

    def step_e2e():
        cap = max(4096, 128 * nf)
        cols = _lib.pinned_empty((ncols, cap), np.float32)
        found = C.c_size_t(0)
        if nf:
            _lib.check(lib.pb_localize(_lib.ptr(hm.array), 0, nf, Y, X, f0, 7, 5000.0, None, 100.0, 1.0, 1.0, 2,
                                       0.001, 100, 0, _lib.ptr(cols), cap, C.byref(found)))
        mine = torch.from_numpy(np.ascontiguousarray(cols[:, : int(found.value)])).to(ctx.dev)
        result["e2e_cols"] = gather(mine).cpu()

    e2e_seconds = ctx.timed(step_e2e)
    hm.free()

    # ---- dominant kernel: identify over the resident block, CUDA events on the launch stream ----
    vp, sz = C.c_void_p, C.c_size_t
    lib.pb_identify_dev.argtypes = [vp, C.c_int, sz, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_double, vp,
                                    vp, vp, vp, vp, sz, vp, vp]
    lib.pb_identify_dev.restype = C.c_int
    cap = max(4096, 512 * nf)
    ifr = torch.empty(cap, dtype=torch.int64, device=ctx.dev); ix = torch.empty_like(ifr); iy = torch.empty_like(ifr)
    ing = torch.empty(cap, dtype=torch.float32, device=ctx.dev)
    icnt = torch.zeros(1, dtype=torch.int64, device=ctx.dev)
    st = torch.cuda.current_stream(ctx.dev)
    ms = []
    for it in range(4):
        icnt.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if nf:
            _lib.check(lib.pb_identify_dev(movie.data_ptr(), 0, nf, Y, X, f0, 7, 5000.0, None, ifr.data_ptr(),
                                           ix.data_ptr(), iy.data_ptr(), ing.data_ptr(), cap, icnt.data_ptr(),
                                           st.cuda_stream))
        e1.record()
        torch.cuda.synchronize(ctx.dev)
        if it:
            ms.append(e0.elapsed_time(e1))
    ident_ms = float(np.mean(ms))
    n_ident = int(icnt.item())
    alg = 2.0 * nf * Y * X + 28.0 * n_ident
    ach = alg / (ident_ms * 1e-3) / 1e9 if ident_ms > 0 else 0.0

    # ---- parity ----
    parity = {}
    if ctx.world > 1:
        # the SAME movie on one GPU: the blocks of all ranks are all-gathered (regenerating them would not
        # be bit-reproducible -- the generator accumulates overlapping emitters with float atomics)
        fmax = max(cb[r + 1] - cb[r] for r in range(ctx.world)) * chunk
        pad = torch.zeros((fmax, Y, X), dtype=torch.int16, device=ctx.dev)
        pad[:nf] = movie
        allb = torch.empty((ctx.world, fmax, Y, X), dtype=torch.int16, device=ctx.dev)
        dist.all_gather_into_tensor(allb.view(torch.uint8).view(-1), pad.view(torch.uint8).view(-1))   # NCCL has no int16
        full = torch.cat([allb[r, : (cb[r + 1] - cb[r]) * chunk] for r in range(ctx.world)])
        del allb, pad
        one = pbd.localize_device(torch, full, 0, CAM, PARAMS, fitting_method="gausslq")
        same = bool(one.shape == table.shape and torch.equal(one.view(torch.int32), table.view(torch.int32)))
        same_e2e = bool(torch.equal(result["e2e_cols"].view(torch.int32), table.cpu().view(torch.int32)))
        parity = {"vs_1gpu_bit_identical": ctx.all_true(same), "e2e_equals_device_run": ctx.all_true(same_e2e)}
        del full
    else:
        import oracle

        sub = movie[:8].cpu().numpy().view(np.uint16)
        ofr, ox, oy, ong = oracle.identify_movie(sub, 5000, 7)
        fr = table[0].view(torch.int32).cpu().numpy()
        sel = fr < 8
        ok = int(sel.sum()) == len(ofr) and np.array_equal(fr[sel], ofr.astype(np.int32))
        parity = {"vs_oracle_identifications_first_8_frames": bool(ok), "n_first_8_frames": int(len(ofr)),
                  "e2e_equals_device_run": bool(torch.equal(result["e2e_cols"].view(torch.int32),
                                                            table.cpu().view(torch.int32)))}
    return {
        "workload": f"configs[2]: {frames}x{Y}x{X} uint16 movie, 60 emitters/frame, box 7, min net gradient 5000, "
                    "gausslq; identify + get_spots + fit + locs_from_fits fused on the GPU",
        "sharding": f"frames in contiguous blocks over {ctx.world} GPU(s); all-gather of the column blocks",
        "seconds": seconds, "e2e_seconds": e2e_seconds, "frames_per_s": frames / seconds,
        "e2e_frames_per_s": frames / e2e_seconds, "n_localizations": n_locs,
        "h2d_bytes_per_rank": nf * Y * X * 2, "d2h_bytes_per_rank": 44 * n_locs // max(ctx.world, 1),
        "kernels_ms": {"identify_kernel": ident_ms},
        "roofline": {"bound": "hbm", "kernel": "identify_kernel (fused local-max + net-gradient)",
                     "achieved": ach, "peak": ctx.peak, "unit": "GB/s", "frac": ach / ctx.peak,
                     "algorithmic_bytes": alg, "traffic": None,
                     "note": "2 B/pixel + 28 B/detection per rank (SURVEY 8d); the kernel is shared-memory / "
                             "issue bound (profiles/r01_summary.md), end to end the stage is PCIe bound"},
        "parity": parity,
    }


# ---------------------------------------------------------------------------------------------
# config 4: render
# ---------------------------------------------------------------------------------------------
def gen_locs_chunk(torch, chunk_index, n, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(42_000 + chunk_index)
    u = torch.rand((4, n), generator=g, device=dev, dtype=torch.float32)
    return (u[0] * 512.0).contiguous(), (u[1] * 512.0).contiguous(), (0.02 + 0.06 * u[2]).contiguous(), \
        (0.02 + 0.06 * u[3]).contiguous()


def stage_render(ctx, n_total=50_000_000, chunk=125_000):
    torch, dist = ctx.torch, ctx.dist
    from picasso_b200 import distributed as pbd

    nchunks = n_total // chunk
    cb = _bounds(nchunks, ctx.world)

    def gen(chunks):
        parts = [gen_locs_chunk(torch, c, chunk, ctx.dev) for c in chunks]
        return tuple(torch.cat([p[k] for p in parts]) for k in range(4))

    x, y, lpx, lpy = gen(range(cb[ctx.rank], cb[ctx.rank + 1]))
    kw = dict(oversampling=20.0, viewport=[(0, 0), (512, 512)], min_blur_width=0.0, blur_method="gaussian")
    d = dist if ctx.world > 1 else None
    res = {}

    def step_dev():
        res["n"], res["band"], res["rows"] = pbd.render_bands_device(d, torch, x, y, lpx, lpy, **kw)

    seconds = ctx.timed(step_dev)
    timings = {}
    ctx.barrier()
    pbd.render_bands_device(d, torch, x, y, lpx, lpy, timings=timings, **kw)
    n_in_view, band, (row0, row1) = res["n"], res["band"], res["rows"]

    # ---- end to end: this rank's share as pageable numpy columns -> its band on the host ----
    import pandas as pd

    host = pd.DataFrame({"x": x.cpu().numpy(), "y": y.cpu().numpy(), "lpx": lpx.cpu().numpy(), "lpy": lpy.cpu().numpy()})
    info = [{"Height": 512, "Width": 512, "Frames": 1, "Pixelsize": 130}]

    def step_e2e():
        res["e2e"] = pbd.render_bands(d, torch, host, info, device=ctx.dev, oversampling=20.0, blur_method="gaussian")

    e2e_seconds = ctx.timed(step_e2e)
    n_e2e, band_e2e, _ = res["e2e"]
    n_mine = int(x.numel())
    del host

    # ---- parity ----
    parity = {}
    bd = band
    if ctx.world > 1:
        fx, fy, flx, fly = gen(range(nchunks))
        n1, full, _ = pbd.render_bands_device(None, torch, fx, fy, flx, fly, **kw)
        ref = full[row0:row1]
        big = ref > 1e-3 * full.max()
        rel = float(((bd - ref).abs()[big] / ref[big]).max().item()) if bool(big.any()) else 0.0
        parity = {"vs_1gpu_n_equal": ctx.all_true(n1 == n_in_view),
                  "vs_1gpu_max_rel_pixel_diff": ctx.max_over_ranks(rel),
                  "vs_1gpu_ok": ctx.all_true(n1 == n_in_view and rel <= 1e-4)}
        del fx, fy, flx, fly, full
    else:
        import oracle

        m = 200_000
        sub = pd.DataFrame({"x": x[:m].cpu().numpy(), "y": y[:m].cpu().numpy(), "lpx": lpx[:m].cpu().numpy(),
                            "lpy": lpy[:m].cpu().numpy()})
        ok_n, oimg = oracle.render(sub, [{"Height": 512, "Width": 512}], oversampling=20.0, blur_method="gaussian")
        n_s, img_s, _ = pbd.render_bands_device(None, torch, x[:m], y[:m], lpx[:m], lpy[:m], **kw)
        img_s = img_s.cpu().numpy()
        big = oimg > 1e-3 * oimg.max()
        rel = float((np.abs(img_s - oimg)[big] / oimg[big]).max())
        parity = {"vs_oracle_200k_sample_n_equal": bool(ok_n == n_s), "vs_oracle_max_rel_pixel_diff": rel,
                  "vs_oracle_ok": bool(ok_n == n_s and rel <= 1e-4)}
    parity["e2e_n_equal"] = ctx.all_true(n_e2e == n_in_view)
    parity["e2e_band_max_abs_diff_vs_device_run"] = ctx.max_over_ranks(
        float(np.abs(band_e2e - bd.cpu().numpy()).max()) if band_e2e.size else 0.0)
    splat_ms = timings.get("splat_ms", 0.0)
    n_recv_bytes = 16.0 * n_total / ctx.world
    alg = n_recv_bytes + 4.0 * (row1 - row0) * 10240
    ach = alg / (splat_ms * 1e-3) / 1e9 if splat_ms > 0 else 0.0
    return {
        "workload": "configs[3]: 50M localizations (x, y ~ U(0, 512), lp ~ U(0.02, 0.08)) -> render gaussian at "
                    "oversampling 20 = 10240 x 10240 float32",
        "sharding": (f"image in {ctx.world} row bands; localizations bucketed by band on the GPU (3 sigma halo), one "
                     "NCCL all-to-all per column, each rank splats and downloads its band" if ctx.world > 1
                     else "single GPU: full image"),
        "seconds": seconds, "e2e_seconds": e2e_seconds, "locs_per_s": n_total / seconds,
        "e2e_locs_per_s": n_total / e2e_seconds, "n_in_view": int(n_in_view),
        "h2d_bytes_per_rank": 16 * n_mine, "d2h_bytes_per_rank": 4 * (row1 - row0) * 10240,
        "phases_ms": {k: ctx.max_over_ranks(v) for k, v in sorted(timings.items())},
        "roofline": {"bound": "hbm", "kernel": "pb_render_band_dev (bin + scan + scatter + render_tiled_kernel)",
                     "achieved": ach, "peak": ctx.peak, "unit": "GB/s", "frac": ach / ctx.peak,
                     "algorithmic_bytes": alg, "traffic": None,
                     "note": "16 B/localization read + one write of the band (SURVEY 8d); the splat is bound by "
                             "shared-memory atomics / issue, not HBM (DESIGN.md 5.5)"},
        "parity": parity,
    }


# ---------------------------------------------------------------------------------------------
# config 5: undrift
# ---------------------------------------------------------------------------------------------
def gen_drift_segment(torch, seg, bounds, n_frames, side, centres, dev, per_frame=1000.0, jitter=0.05, lp=0.05,
                      amp_x=1.0, amp_y=0.7):
    """Localizations of one segment (config 5, SURVEY.md 8d: clusters + the analytic drift of the
    reference's tests/test_undrift.py scaled to n_frames; same model as testing.synthetic_drift_locs)."""
    g = torch.Generator(device=dev)
    g.manual_seed(53_000 + seg)
    f_lo, f_hi = int(bounds[seg]), int(bounds[seg + 1])
    n = int((f_hi - f_lo) * per_frame)
    frame = torch.sort(torch.randint(f_lo, max(f_hi, f_lo + 1), (n,), generator=g, device=dev)).values
    which = torch.randint(0, centres.shape[0], (n,), generator=g, device=dev)
    t = frame.double()
    dx = amp_x * torch.sin(2 * math.pi * t / (n_frames / 2.0))
    dy = amp_y * (t / n_frames - 0.5)
    nz = torch.randn((2, n), generator=g, device=dev, dtype=torch.float64) * jitter
    x = (centres[which, 0] + dx + nz[0]).float()
    y = (centres[which, 1] + dy + nz[1]).float()
    lpv = torch.full((n,), lp, dtype=torch.float32, device=dev)
    return frame, x, y, lpv


def stage_undrift(ctx, n_frames=20000, side=4096, segmentation=100, n_clusters=2000):
    torch, dist = ctx.torch, ctx.dist
    import pandas as pd
    from scipy import interpolate

    from picasso_b200 import distributed as pbd, lib as pblib

    n_seg = int(np.round(n_frames / segmentation))
    bounds = np.linspace(0, n_frames - 1, n_seg + 1, dtype=np.uint32)
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(5)
    centres = 6 + (side - 12) * torch.rand((n_clusters, 2), generator=g, device=ctx.dev, dtype=torch.float64)
    sb = pbd.segment_shards(n_seg, ctx.world)

    def gen(segs):
        parts = [gen_drift_segment(torch, s, bounds, n_frames, side, centres, ctx.dev) for s in segs]
        if not parts:
            z = torch.empty(0, dtype=torch.float32, device=ctx.dev)
            return torch.empty(0, dtype=torch.int64, device=ctx.dev), z, z, z, np.zeros(1, np.int64)
        start = np.concatenate([[0], np.cumsum([int(p[0].numel()) for p in parts])]).astype(np.int64)
        return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts]),
                torch.cat([p[3] for p in parts]), start)

    frame, x, y, lp, seg_start = gen(range(sb[ctx.rank], sb[ctx.rank + 1]))
    d = dist if ctx.world > 1 else None
    res = {}

    def drift_from(sy, sx):
        shifts_x = np.zeros((n_seg, n_seg)); shifts_y = np.zeros((n_seg, n_seg))
        ai, aj = np.triu_indices(n_seg, 1)
        shifts_y[ai, aj] = sy; shifts_x[ai, aj] = sx
        shift_y, shift_x = pblib.minimize_shifts(shifts_x, shifts_y)
        t = (bounds[1:] + bounds[:-1]) / 2
        ti = np.arange(n_frames)
        return (interpolate.InterpolatedUnivariateSpline(t, shift_x, k=3)(ti),
                interpolate.InterpolatedUnivariateSpline(t, shift_y, k=3)(ti))

    def step_dev():
        sy, sx = pbd.undrift_shifts_device(d, torch, seg_start, x, y, lp, lp, n_seg, side, side)
        res["sy"], res["sx"] = sy, sx
        res["drift"] = drift_from(sy, sx)

    seconds = ctx.timed(step_dev, warmup=1)
    timings = {}
    ctx.barrier()
    pbd.undrift_shifts_device(d, torch, seg_start, x, y, lp, lp, n_seg, side, side, timings=timings)
    sy, sx = res["sy"], res["sx"]

    # ---- end to end: this rank's rows as a host DataFrame -> drift + its undrifted rows ----
    info = [{"Height": side, "Width": side, "Frames": n_frames, "Pixelsize": 130}]
    host = pd.DataFrame({"frame": frame.cpu().numpy().astype(np.uint32), "x": x.cpu().numpy(), "y": y.cpu().numpy(),
                         "lpx": lp.cpu().numpy(), "lpy": lp.cpu().numpy()})

    def step_e2e():
        res["e2e"] = pbd.undrift_segments_sharded(d, torch, host, info, segmentation, device=ctx.dev)

    e2e_seconds = ctx.timed(step_e2e, warmup=1)
    n_mine = len(host)
    n_total = int(ctx.sum_over_ranks(n_mine))
    del host

    # ---- parity ----
    ti = np.arange(n_frames)
    truth_x = 1.0 * np.sin(2 * np.pi * ti / (n_frames / 2.0))
    truth_y = 0.7 * (ti / n_frames - 0.5)
    dxv, dyv = res["drift"]
    err = [float(np.abs((dxv - dxv.mean()) - (truth_x - truth_x.mean())).max()),
           float(np.abs((dyv - dyv.mean()) - (truth_y - truth_y.mean())).max())]
    parity = {"max_abs_drift_error_px_vs_injected": err, "injected_drift_recovered": bool(max(err) < 0.01)}
    e2e_drift = res["e2e"][0]
    parity["e2e_drift_max_abs_diff_vs_device_run"] = ctx.max_over_ranks(
        float(max(np.abs(e2e_drift["x"].to_numpy() - dxv).max(), np.abs(e2e_drift["y"].to_numpy() - dyv).max())))
    if ctx.world > 1:
        _, fx, fy, flp, fstart = gen(range(n_seg))
        s1y, s1x = pbd.undrift_shifts_device(None, torch, fstart, fx, fy, flp, flp, n_seg, side, side)
        diff = float(max(np.abs(s1y - sy).max(), np.abs(s1x - sx).max()))
        parity["vs_1gpu_max_abs_pair_shift_diff_px"] = ctx.max_over_ranks(diff)
        parity["vs_1gpu_ok"] = ctx.all_true(diff <= 1e-4)
        del fx, fy, flp
    n_pairs = n_seg * (n_seg - 1) // 2
    pairs_ms = timings.get("pairs_ms", 0.0)
    my_pairs = len(pbd.my_tile_pairs(n_seg, 1, ctx.rank, ctx.world)[0])
    spec_bytes = side * (side // 2 + 1) * 8
    alg = 2.0 * spec_bytes * my_pairs
    ach = alg / (pairs_ms * 1e-3) / 1e9 if pairs_ms > 0 else 0.0
    return {
        "workload": f"configs[4]: postprocess.undrift, {n_total / 1e6:.0f}M localizations, {n_frames} frames of "
                    f"{side}x{side}, segmentation {segmentation} -> {n_seg} segments, {n_pairs} pairs",
        "sharding": (f"segments in contiguous blocks over {ctx.world} GPUs for render + R2C; one in-place NCCL "
                     "all-gather of the half-spectra; pairs by whole L2 tiles incl. peak fits; all-gather of 2 "
                     "float64 shifts per pair" if ctx.world > 1 else "single GPU"),
        "seconds": seconds, "e2e_seconds": e2e_seconds, "pairs_per_s": n_pairs / seconds,
        "n_localizations": n_total, "h2d_bytes_per_rank": 16 * n_mine, "d2h_bytes_per_rank": 64 * my_pairs,
        "phases_ms": {k: ctx.max_over_ranks(v) for k, v in sorted(timings.items())},
        "roofline": {"bound": "hbm", "kernel": "pair stage (rcc_fft_rows_async_kernel + rcc_cols_gemm_kernel)",
                     "achieved": ach, "peak": ctx.peak, "unit": "GB/s", "frac": ach / ctx.peak,
                     "algorithmic_bytes": alg, "traffic": None,
                     "note": "algorithmic = 2 half-spectra (2 x 67 MB) read per pair (SURVEY 8d); the L2 pair "
                             "tiling serves most of those reads from L2 (ncu: 8.9 MB of DRAM traffic per pair, "
                             "profiles/r01_summary.md), so the achieved figure may exceed the HBM peak"},
        "parity": parity,
    }


def run_stages(torch, dist, rank, world, dev, peak_gbs, which=("localize", "render", "undrift"), repeats=3):
    ctx = Ctx(torch, dist, rank, world, dev, peak_gbs, repeats)
    out = {}
    table = {"localize": ("localize_config3", stage_localize), "render": ("render_config4", stage_render),
             "undrift": ("undrift_config5", stage_undrift)}
    for w in which:
        name, fn = table[w]
        t0 = time.perf_counter()
        try:
            out[name] = fn(ctx)
            out[name]["stage_wall_s"] = time.perf_counter() - t0
        except Exception as exc:      # noqa: BLE001 -- a failing stage must not lose the headline line
            import traceback

            traceback.print_exc()
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        ctx.barrier()
    return out
