"""Multi-GPU plumbing over ``torch.distributed`` (one process per GPU; NCCL on B200,
gloo in the CPU tests).  SURVEY.md section 8e: every stage of the hot path shards without
a data-path collective; the only exchanges are the gather of results.

* fits: spots shard by contiguous index range; ONE all-gather of the packed per-rank
  output buffer ``[thetas 6n | crlbs 6n | logliks n | iterations n]``.
* identify: frames shard by contiguous range; variable-length results are exchanged by an
  all-gather of counts followed by a padded all-gather.
* render: the IMAGE shards by row bands (SURVEY.md 8e option B): every rank buckets its index
  share of the localisations by destination band (3-sigma halo) on its GPU, one NCCL all-to-all
  moves the records, every rank splats and downloads only its band -- no 419 MB image reduction.
  (``render_sharded`` keeps the simpler index shard + image all-reduce.)
* fused localize: frames shard by contiguous block, the finished column blocks are gathered.
* undrift: SEGMENTS shard for rendering + forward R2C, one NCCL all-gather leaves every rank with
  all half-spectra (13.4 GB at 200 x 4096^2), the n(n-1)/2 pairs shard by whole L2 pair tiles
  incl. their peak fits; two float64 shifts per pair are exchanged.  (``undrift_sharded`` keeps the
  variant without a spectra exchange: every rank renders and transforms all segments.)
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int):
    """Contiguous ranges [r*n/world, (r+1)*n/world) -- equal to within one element."""
    return [(n * r) // world for r in range(world + 1)]


def my_shard(n: int, rank: int, world: int):
    b = shard_bounds(n, world)
    return b[rank], b[rank + 1]


def pack_fit_outputs(torch, thetas, crlbs, logliks, iterations):
    """[thetas 6n | crlbs 6n | logliks n | iterations n] as one flat float32 tensor
    (iterations bit-cast)."""
    return torch.cat([thetas.reshape(-1), crlbs.reshape(-1), logliks.reshape(-1),
                      iterations.view(torch.float32).reshape(-1)])


def unpack_fit_outputs(torch, flat, n):
    th = flat[: 6 * n].reshape(n, 6)
    cr = flat[6 * n: 12 * n].reshape(n, 6)
    ll = flat[12 * n: 13 * n]
    it = flat[13 * n: 14 * n].view(torch.int32)
    return th, cr, ll, it


def all_gather_fit_outputs(dist, torch, local_flat, counts):
    """Gather every rank's packed outputs (ranks may differ by one spot: pad to the max).
    Returns the list of per-rank (thetas, crlbs, logliks, iterations)."""
    world = dist.get_world_size()
    nmax = max(counts)
    padded = torch.zeros(14 * nmax, dtype=torch.float32, device=local_flat.device)
    n_mine = counts[dist.get_rank()]
    # re-pack with the padded strides so every rank's block has the same layout
    th, cr, ll, it = unpack_fit_outputs(torch, local_flat, n_mine)
    padded[: 6 * n_mine] = th.reshape(-1)
    padded[6 * nmax: 6 * nmax + 6 * n_mine] = cr.reshape(-1)
    padded[12 * nmax: 12 * nmax + n_mine] = ll
    padded[13 * nmax: 13 * nmax + n_mine] = it.view(torch.float32)
    out = torch.empty(14 * nmax * world, dtype=torch.float32, device=local_flat.device)
    dist.all_gather_into_tensor(out, padded)
    res = []
    for r in range(world):
        blk = out[r * 14 * nmax: (r + 1) * 14 * nmax]
        n_r = counts[r]
        res.append((blk[: 6 * nmax].reshape(nmax, 6)[:n_r],
                    blk[6 * nmax: 12 * nmax].reshape(nmax, 6)[:n_r],
                    blk[12 * nmax: 13 * nmax][:n_r],
                    blk[13 * nmax: 14 * nmax].view(torch.int32)[:n_r]))
    return res


def sharded_fit(dist, torch, spots, fit_fn, device="cpu"):
    """Fit ``spots`` (N, b, b) with this rank's share computed by ``fit_fn(spots_shard) ->
    (thetas, crlbs, logliks, iterations)`` and return the full result on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    N = len(spots)
    bounds = shard_bounds(N, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    th, cr, ll, it = fit_fn(spots[lo:hi])
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    flat = pack_fit_outputs(torch, t(th, torch.float32), t(cr, torch.float32),
                            t(ll, torch.float32), t(it, torch.int32))
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    parts = all_gather_fit_outputs(dist, torch, flat, counts)
    cat = lambda k: torch.cat([p[k] for p in parts]).cpu().numpy()
    return cat(0), cat(1), cat(2), cat(3)


def all_gather_variable(dist, torch, arrays, device="cpu"):
    """All-gather a tuple of equally long 1-D numpy arrays whose length differs per rank
    (identify results): counts first, then a padded gather per array."""
    world = dist.get_world_size()
    n = len(arrays[0])
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    outs = []
    for a in arrays:
        a = np.ascontiguousarray(a)
        pad = torch.zeros(nmax, dtype=torch.from_numpy(a[:0].copy()).dtype, device=device)
        pad[:n] = torch.from_numpy(a).to(device)
        buf = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(buf, pad)
        outs.append(np.concatenate([b[:c].cpu().numpy() for b, c in zip(buf, counts)]))
    return tuple(outs)


def all_reduce_image(dist, torch, image, device="cpu"):
    """Sum per-rank partial renders (render sharded by localisation index)."""
    t = torch.from_numpy(np.ascontiguousarray(image)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


# ---- higher-level sharded stages ---------------------------------------------------------
def gather_column_blocks(dist, torch, cols, device="cpu"):
    """All-gather (ncols, n_r) float32 blocks whose n_r differs per rank and concatenate them in
    rank order (fused localize: ranks own consecutive frame blocks, so the result stays sorted)."""
    world = dist.get_world_size()
    cols = np.ascontiguousarray(cols, dtype=np.float32)
    ncols, n = cols.shape
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    pad = torch.zeros((ncols, nmax), dtype=torch.float32, device=device)
    pad[:, :n] = torch.from_numpy(cols).to(device)
    out = torch.empty((world, ncols, nmax), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
    out = out.cpu().numpy()
    return np.concatenate([out[r][:, :counts[r]] for r in range(world)], axis=1)


def my_pairs(n_seg: int, rank: int, world: int):
    """This rank's share of the i < j segment pairs (round-robin over the reference's pair order)."""
    pi, pj = np.triu_indices(n_seg, 1)
    return pi[rank::world].astype(np.int32), pj[rank::world].astype(np.int32)


def gather_pair_windows(dist, torch, win_mine, n_seg, device="cpu"):
    """All-gather the per-rank correlation windows (round-robin pair shards) back into the
    reference's pair order: returns (n_pairs, H, W) float32 on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_pairs = n_seg * (n_seg - 1) // 2
    win_mine = np.ascontiguousarray(win_mine, dtype=np.float32)
    H, W = win_mine.shape[1:]
    counts = [len(range(r, n_pairs, world)) for r in range(world)]
    nmax = max(max(counts), 1)
    pad = torch.zeros((nmax, H, W), dtype=torch.float32, device=device)
    pad[: counts[rank]] = torch.from_numpy(win_mine).to(device)
    out = torch.empty((world, nmax, H, W), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
    out = out.cpu().numpy()
    full = np.zeros((n_pairs, H, W), np.float32)
    for r in range(world):
        full[r::world] = out[r][: counts[r]]
    return full


def localize_sharded(dist, torch, movie, camera_info, parameters, *, fitting_method="gausslq", eps=0.001,
                     max_it=100, mle_method="sigmaxy", roi=None, frame_bounds=None, device="cuda"):
    """``localize.localize`` with the frames sharded over the ranks (contiguous blocks): every rank
    runs the fused movie -> table pass on its block; the column blocks are gathered so that every
    rank returns the full localization table."""
    from . import localize as pbl

    rank, world = dist.get_rank(), dist.get_world_size()
    N = len(movie)
    lo, hi = pbl._frame_range(N, frame_bounds)
    hi = min(hi, N - 1)
    b = shard_bounds(max(hi - lo + 1, 0), world)
    mine = (lo + b[rank], lo + b[rank + 1] - 1)
    fit = pbl._FIT_IDS[(fitting_method, mle_method if fitting_method == "gaussmle" else None)]
    names = pbl.LOCS_COLUMNS_MLE if fit <= 1 else pbl.LOCS_COLUMNS_LQ
    if mine[1] >= mine[0]:
        cols = pbl._localize_fused_columns(movie, parameters["Min. Net Gradient"], parameters["Box Size"],
                                           camera_info, fit, eps, max_it, roi=roi, frame_bounds=mine)
    else:
        cols = np.zeros((len(names), 0), np.float32)
    full = gather_column_blocks(dist, torch, cols, device=device)
    return pbl._columns_to_locs(full, names)


def render_sharded(dist, torch, locs, info, device="cuda", **kwargs):
    """``render.render`` with the localizations sharded by index: every rank splats its share into
    a device image, the partial images are summed with ONE all-reduce on the GPUs (NVLink) and
    downloaded once.  Returns (n, image) on every rank."""
    from . import render as pbr

    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = my_shard(len(locs), rank, world)
    cnt, image = pbr.render_to_device(torch, locs.iloc[lo:hi], info, device=device, **kwargs)
    dist.all_reduce(image, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return int(cnt.item()), pbr.image_to_host(torch, image)


def gather_pair_values(dist, torch, vals_mine, n_seg, device="cpu"):
    """All-gather per-pair float64 rows (round-robin pair shards) into the reference's pair order:
    ``vals_mine`` (n_mine, k) -> (n_pairs, k) on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_pairs = n_seg * (n_seg - 1) // 2
    vals_mine = np.ascontiguousarray(vals_mine, dtype=np.float64)
    k = vals_mine.shape[1]
    counts = [len(range(r, n_pairs, world)) for r in range(world)]
    nmax = max(max(counts), 1)
    pad = torch.zeros((nmax, k), dtype=torch.float64, device=device)
    pad[: counts[rank]] = torch.from_numpy(vals_mine).to(device)
    out = torch.empty((world, nmax, k), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
    out = out.cpu().numpy()
    full = np.zeros((n_pairs, k))
    for r in range(world):
        full[r::world] = out[r][: counts[r]]
    return full


def undrift_sharded(dist, torch, locs, info, segmentation, device="cuda"):
    """``postprocess.undrift`` with the segment pairs sharded round-robin over the ranks: every
    rank renders and transforms all segments, correlates its pairs and fits their peaks on its GPU;
    only two float64 shifts per pair are exchanged before ``minimize_shifts``."""
    from . import imageprocess, lib, postprocess

    rank, world = dist.get_rank(), dist.get_world_size()

    def shifts(locs_, info_, bounds, min_blur_width, max_shift, callback):
        n_seg = len(bounds) - 1
        pi, pj = my_pairs(n_seg, rank, world)
        sy, sx = imageprocess._shifts_of_locs(locs_, info_, bounds, min_blur_width, max_shift, pairs=(pi, pj))
        full = gather_pair_values(dist, torch, np.stack([sy, sx], 1), n_seg, device=device)
        shifts_x = np.zeros((n_seg, n_seg))
        shifts_y = np.zeros((n_seg, n_seg))
        ai, aj = np.triu_indices(n_seg, 1)
        shifts_y[ai, aj] = full[:, 0]
        shifts_x[ai, aj] = full[:, 1]
        return lib.minimize_shifts(shifts_x, shifts_y)

    return postprocess.undrift(locs, info, segmentation, display=False,
                               segmentation_callback=lambda i: None, rcc_callback=lambda i: None,
                               _shifts_fn=shifts)


class PeerGather:
    """All-gather of equally sized per-rank device blocks WITHOUT a kernel: every rank owns a
    gather buffer allocated by the library (``pb_dev_alloc``), shares it with the other ranks of the
    box through CUDA IPC handles (exchanged once over ``torch.distributed``) and receives the
    blocks by copy-engine peer writes (``pb_copy_d2d_async`` over NVLink / NVSwitch).  Unlike an
    NCCL all-gather no SM is taken from the fit kernel that runs at the same time, and nothing
    blocks the host: ``gather_async`` returns CUDA events, ``wait`` makes a stream wait for them.

    The block of rank r lands at byte offset ``r * block_bytes`` of every rank's buffer."""

    def __init__(self, dist, torch, block_bytes: int, device, n_streams: int = 4):
        import ctypes as C

        from . import _lib

        self.dist, self.torch, self.C, self._lib = dist, torch, C, _lib
        self.l = l = _lib.load()
        vp, sz = C.c_void_p, C.c_size_t
        l.pb_dev_alloc.argtypes = [C.POINTER(vp), sz]
        l.pb_dev_free.argtypes = [vp]
        l.pb_ipc_export.argtypes = [vp, vp]
        l.pb_ipc_open.argtypes = [vp, C.POINTER(vp)]
        l.pb_ipc_close.argtypes = [vp]
        l.pb_copy_d2d_async.argtypes = [vp, vp, sz, vp]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.block_bytes = int(block_bytes)
        self.device = device
        buf = vp()
        _lib.check(l.pb_dev_alloc(C.byref(buf), self.block_bytes * self.world))
        self.buf = buf.value
        hb = l.pb_ipc_handle_bytes()
        handle = np.zeros(hb, np.uint8)
        _lib.check(l.pb_ipc_export(self.buf, handle.ctypes.data))
        mine = torch.from_numpy(handle).to(device)
        everyone = torch.empty(hb * self.world, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(everyone, mine)
        handles = everyone.cpu().numpy().reshape(self.world, hb)
        self.peers = []
        for r in range(self.world):
            if r == self.rank:
                self.peers.append(self.buf)
            else:
                p = vp()
                h = np.ascontiguousarray(handles[r])
                _lib.check(l.pb_ipc_open(h.ctypes.data, C.byref(p)))
                self.peers.append(p.value)
        self.streams = [torch.cuda.Stream(device) for _ in range(max(1, min(n_streams, self.world)))]
        dist.barrier()

    def gather_async(self, block, stream=None, offset_bytes: int = 0, after_event=None):
        """Enqueue the copies of ``block`` (a contiguous device tensor; the whole block of this rank,
        or the part of it that starts ``offset_bytes`` into the block) into every rank's buffer,
        ordered after the work already enqueued on ``stream`` -- or after ``after_event`` (e.g. the
        fit's phase event: the thetas are final before the CRLB kernel has run); returns the events
        that complete when the data has left this rank."""
        torch = self.torch
        stream = stream or torch.cuda.current_stream(self.device)
        nbytes = block.numel() * block.element_size()
        assert block.is_contiguous() and offset_bytes + nbytes <= self.block_bytes
        if after_event is not None:
            ready = after_event
        else:
            ready = torch.cuda.Event()
            ready.record(stream)
        for k in range(self.world):
            r = (self.rank + 1 + k) % self.world           # start with the neighbour: spread the links
            st = self.streams[k % len(self.streams)]
            st.wait_event(ready)
            self._lib.check(self.l.pb_copy_d2d_async(
                self.peers[r] + self.rank * self.block_bytes + offset_bytes, block.data_ptr(), nbytes,
                st.cuda_stream))
        done = []
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            done.append(ev)
        return done

    @staticmethod
    def wait(events, stream):
        """Make ``stream`` wait (on the device, the host does not block) for a gather's events."""
        for ev in events or ():
            stream.wait_event(ev)

    def finish(self):
        """All outgoing copies of this rank done, then a barrier: every rank's buffer is complete."""
        for st in self.streams:
            st.synchronize()
        self.dist.barrier()

    def to_tensor(self, dtype):
        """Copy of this rank's gather buffer as a torch tensor (for checks / consumers)."""
        torch = self.torch
        out = torch.empty(self.block_bytes * self.world // torch.empty(0, dtype=dtype).element_size(),
                          dtype=dtype, device=self.device)
        st = torch.cuda.current_stream(self.device)
        self._lib.check(self.l.pb_copy_d2d_async(out.data_ptr(), self.buf, self.block_bytes * self.world,
                                                 st.cuda_stream))
        st.synchronize()
        return out

    def close(self):
        if self.buf is None:
            return
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.l.pb_ipc_close(p)
        self.dist.barrier()                  # nobody frees while a peer still has the mapping open
        self.l.pb_dev_free(self.buf)
        self.buf = None


# ---- scalable sharding of render and undrift (device-resident building blocks) -------------------
def band_rows(n_pixel_y: int, world: int, tile: int = 64):
    """Row bands of an n_pixel_y-row image for ``world`` ranks: world + 1 boundaries, interior ones
    aligned to the 64-row render tile (a band may be empty when the image has fewer tiles than
    ranks)."""
    rows = [min(n_pixel_y, ((n_pixel_y * r) // world + tile // 2) // tile * tile) for r in range(world)]
    rows[0] = 0
    for r in range(1, world):
        rows[r] = max(rows[r], rows[r - 1])
    return rows + [n_pixel_y]


def exchange_variable(dist, torch, send, in_splits, all_splits):
    """All-to-all of a 1-D tensor with per-destination split sizes ``in_splits``; ``all_splits`` is
    the (world, world) matrix of everybody's split sizes (row = source).  Returns the received
    tensor (source-major).  One ``all_to_all_single`` (NCCL on the GPUs, gloo in the CPU tests)."""
    rank = dist.get_rank()
    out_splits = [int(all_splits[src][rank]) for src in range(dist.get_world_size())]
    recv = torch.empty(sum(out_splits), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=out_splits,
                           input_split_sizes=[int(v) for v in in_splits])
    return recv


def all_gather_counts(dist, torch, counts, device):
    """(world, world) int64 matrix of every rank's per-destination counts."""
    world = dist.get_world_size()
    mine = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=device)
    full = torch.empty(world * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(full, mine)
    return full.cpu().numpy().reshape(world, world)


def _render_decl(l):
    if getattr(l, "_band_declared", False):
        return
    import ctypes as C

    vp, i32, f64, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    l.pb_render_band_dev.argtypes = [sz, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, i32, vp, i32, i32,
                                     i32, i32, vp, vp, sz, vp]
    l.pb_render_band_dev.restype = i32
    l.pb_render_band_count_dev.argtypes = [sz, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, i32, i32, i32,
                                           i32, vp, vp, vp]
    l.pb_render_band_count_dev.restype = i32
    l.pb_render_band_scatter_dev.argtypes = [sz, vp, vp, vp, vp, f64, f64, f64, f64, f64, f64, i32, i32, i32,
                                             i32, vp, vp, vp, vp, vp]
    l.pb_render_band_scatter_dev.restype = i32
    l.pb_render_unpack_records_dev.argtypes = [sz, vp, vp, vp, vp, vp, vp]
    l.pb_render_unpack_records_dev.restype = i32
    l.pb_render_workspace_bytes.restype = C.c_size_t
    l.pb_render_workspace_bytes.argtypes = [sz, i32, i32]
    l._band_declared = True


def render_bands_device(dist, torch, x, y, lpx, lpy, *, oversampling, viewport, min_blur_width=0.0,
                        blur_method="gaussian", timings=None):
    """Multi-GPU ``render.render`` by image row bands.  ``x, y, lpx, lpy`` are float32 CUDA tensors
    holding THIS rank's index share of the localizations.  Every rank

    1. buckets its localizations by destination band on its GPU (``pb_render_band_count_dev`` /
       ``pb_render_band_scatter_dev``: a localization goes to every band its 3-sigma window reaches),
    2. exchanges the (x, y, lpx, lpy) records with ONE all-to-all (NCCL over NVLink),
    3. splats what it received into its own band (``pb_render_band_dev``, windows clipped to the
       band) -- no image reduction, and every rank downloads only its band.

    Returns ``(n_total, band_image (rows, n_pixel_x) CUDA tensor, (row0, row1))``; the bands
    concatenate to the single-GPU image (up to float32 summation order) and ``n_total`` is the
    reference's ``n``.  ``dist=None`` or world size 1 renders the full image locally."""
    import ctypes as C

    from . import _lib
    from .render import _MODES

    if blur_method not in _MODES:
        raise Exception("blur_method not understood.")
    mode = _MODES[blur_method]
    l = _lib.load()
    _render_decl(l)
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    dev = x.device
    st = torch.cuda.current_stream(dev).cuda_stream
    (y_min, x_min), (y_max, x_max) = viewport
    npy = int(np.ceil(oversampling * (y_max - y_min)))
    npx = int(np.ceil(oversampling * (x_max - x_min)))
    args = (float(oversampling), float(y_min), float(x_min), float(y_max), float(x_max),
            float(min_blur_width), mode)
    n = int(x.numel())
    p = lambda t: t.data_ptr() if t is not None and t.numel() else None     # noqa: E731
    rows = band_rows(npy, world)
    ev = (lambda: torch.cuda.Event(enable_timing=True)) if timings is not None else None
    marks = []

    def mark(name):
        if ev is not None:
            e = ev(); e.record(); marks.append((name, e))

    mark("start")
    if world > 1:
        rows_c = (C.c_int * (world + 1))(*rows)
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        _lib.check(l.pb_render_band_count_dev(n, p(x), p(y), p(lpx), p(lpy), *args, npy, npx, world, rows_c,
                                              counts.data_ptr(), st))
        all_counts = all_gather_counts(dist, torch, counts.cpu().numpy(), dev)     # one sync
        mine = all_counts[rank]
        offsets = torch.as_tensor(np.concatenate([[0], np.cumsum(mine)[:-1]]).astype(np.int64), device=dev)
        cursor = torch.zeros(world, dtype=torch.int64, device=dev)
        total = int(mine.sum())
        send = torch.empty((max(total, 1), 4), dtype=torch.float32, device=dev)      # (x, y, lpx, lpy) records
        _lib.check(l.pb_render_band_scatter_dev(n, p(x), p(y), p(lpx), p(lpy), *args, npy, npx, world, rows_c,
                                                offsets.data_ptr(), cursor.data_ptr(), send.data_ptr(), st))
        mark("bucket")
        recv = exchange_variable(dist, torch, send.view(-1)[: 4 * total], 4 * mine, 4 * all_counts)   # ONE all-to-all
        mark("exchange")
        n = int(recv.numel()) // 4
        cols = torch.empty((4, max(n, 1)), dtype=torch.float32, device=dev)
        _lib.check(l.pb_render_unpack_records_dev(n, recv.data_ptr() if n else None, cols[0].data_ptr(),
                                                  cols[1].data_ptr(), cols[2].data_ptr(), cols[3].data_ptr(), st))
        x, y = cols[0, :n], cols[1, :n]
        lpx, lpy = (cols[2, :n], cols[3, :n]) if mode else (None, None)
    row0, row1 = rows[rank], rows[rank + 1]
    image = torch.empty((row1 - row0, npx), dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    wsb = int(l.pb_render_workspace_bytes(n, max(row1 - row0, 1), npx)) if mode else 0
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    _lib.check(l.pb_render_band_dev(n, p(x), p(y), p(lpx), p(lpy), *args, p(image), npy, npx, row0, row1 - row0,
                                    count.data_ptr(), ws.data_ptr() if mode else None, wsb, st))
    mark("splat")
    if world > 1:
        dist.all_reduce(count, op=dist.ReduceOp.SUM)
    n_total = int(count.item())
    if timings is not None:
        torch.cuda.synchronize(dev)
        for (na, a), (nb, b) in zip(marks[:-1], marks[1:]):
            timings[nb + "_ms"] = a.elapsed_time(b)
    return n_total, image, (row0, row1)


def render_bands(dist, torch, locs, info, device="cuda", **kwargs):
    """``render.render`` for a table sharded by index over the ranks: ``locs`` is THIS rank's share
    (host DataFrame).  Uploads it, renders by row bands (``render_bands_device``) and downloads this
    rank's band into page-locked memory.  Returns ``(n_total, band ndarray, (row0, row1))``;
    ``gather_bands`` assembles the full image where one is wanted."""
    from . import _lib, lib as pblib
    from .render import image_to_host

    pixelsize = pblib.get_from_metadata(info, "Pixelsize", raise_error=True)
    disp = kwargs.pop("disp_px_size", None)
    oversampling = kwargs.pop("oversampling", 1.0)
    if disp is not None:
        oversampling = pixelsize / disp
    viewport = kwargs.pop("viewport", None) or [(0, 0), (info[0]["Height"], info[0]["Width"])]
    blur = kwargs.pop("blur_method", None)
    up = lambda c: _to_device(torch, np.ascontiguousarray(locs[c], dtype=np.float32), device)   # noqa: E731
    x, y = up("x"), up("y")
    lpx, lpy = (up("lpx"), up("lpy")) if blur is not None else (None, None)
    n, band, rr = render_bands_device(dist, torch, x, y, lpx, lpy, oversampling=oversampling, viewport=viewport,
                                      min_blur_width=kwargs.pop("min_blur_width", 0.0), blur_method=blur)
    return n, image_to_host(torch, band), rr


def gather_bands(dist, torch, band, n_pixel_y, device="cpu"):
    """All-gather the row bands of ``render_bands`` into the full image on every rank."""
    world = dist.get_world_size()
    rows = band_rows(n_pixel_y, world)
    npx = band.shape[1]
    hmax = max(max(rows[r + 1] - rows[r] for r in range(world)), 1)
    pad = torch.zeros((hmax, npx), dtype=torch.float32, device=device)
    pad[: band.shape[0]] = torch.as_tensor(band, device=device)
    out = torch.empty((world, hmax, npx), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
    out = out.cpu().numpy()
    return np.concatenate([out[r][: rows[r + 1] - rows[r]] for r in range(world)], axis=0)


def _to_device(torch, a, device):
    """Host numpy array -> device tensor through the threaded pinned staging (pb_copy_h2d)."""
    from . import _lib

    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, device=device)
    if a.size:
        l = _lib.load()
        st = torch.cuda.current_stream(t.device)
        _lib.check(l.pb_copy_h2d(t.data_ptr(), a.ctypes.data, a.nbytes, st.cuda_stream))
        st.synchronize()
    return t


# ---- undrift: segments sharded for render + R2C, spectra all-gathered, pairs sharded by L2 tile ----
def segment_shards(n_seg: int, world: int):
    """Contiguous segment ranges per rank (world + 1 boundaries)."""
    return shard_bounds(n_seg, world)


def tile_sorted_pairs(n_seg: int, tile_segments: int):
    """All i < j pairs ordered by L2 pair tile (i // TS, j // TS) (stable: the reference's pair order
    inside a tile) and, for each, its index in the reference's pair order."""
    pi, pj = np.triu_indices(n_seg, 1)
    key = (pi // tile_segments).astype(np.int64) * (n_seg + 1) + (pj // tile_segments)
    order = np.argsort(key, kind="stable")
    return pi[order].astype(np.int32), pj[order].astype(np.int32), order


def my_tile_pairs(n_seg: int, tile_segments: int, rank: int, world: int):
    """This rank's contiguous share of the tile-sorted pair list: whole tiles keep their 2 * TS
    spectrum slabs L2 resident (a round-robin split would load every tile on every rank).
    Returns (pair_i, pair_j, index in the reference's pair order)."""
    pi, pj, order = tile_sorted_pairs(n_seg, tile_segments)
    b = shard_bounds(len(pi), world)
    sl = slice(b[rank], b[rank + 1])
    return pi[sl], pj[sl], order[sl]


def _rcc_decl(l):
    if getattr(l, "_rccdev_declared", False):
        return
    import ctypes as C

    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    l.pb_rcc_spectra_dev.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    l.pb_rcc_spectra_dev.restype = i32
    l.pb_rcc_windows_dev.argtypes = [i32, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, vp, sz, vp]
    l.pb_rcc_windows_dev.restype = i32
    l.pb_rcc_peakfit_dev.argtypes = [i32, vp, i32, i32, vp, vp]
    l.pb_rcc_peakfit_dev.restype = i32
    l.pb_rcc_tile_segments.argtypes = [i32]
    l.pb_rcc_tile_segments.restype = i32
    l.pb_render_dev.argtypes = [sz, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_double, i32, vp, i32, i32, vp, vp, sz, vp]
    l.pb_render_dev.restype = i32
    l.pb_render_workspace_bytes.restype = C.c_size_t
    l.pb_render_workspace_bytes.argtypes = [sz, i32, i32]
    l._rccdev_declared = True


def undrift_shifts_device(dist, torch, seg_start, x, y, lpx, lpy, n_seg, Y, X, *, min_blur_width=1.0,
                          max_shift=32, timings=None):
    """The device part of ``postprocess.undrift`` on ``world`` GPUs.  ``x, y, lpx, lpy`` are float32
    CUDA tensors with the localizations of THIS rank's segments (``segment_shards``), grouped by
    segment: local segment k owns ``[seg_start[k], seg_start[k + 1])``.

    1. every rank renders its segments (``pb_render_dev``) and transforms them (one batched R2C,
       ``pb_rcc_spectra_dev``) straight into its slice of the full spectra buffer;
    2. ONE in-place NCCL all-gather leaves all ranks with all half-spectra;
    3. every rank correlates its share of the pairs -- whole L2 pair tiles (``my_tile_pairs``) --
       with the pruned inverse transform (``pb_rcc_windows_dev``) and fits their peaks on the device
       (``pb_rcc_peakfit_dev``); fits the device solver does not settle are re-fitted on the host
       exactly like the reference;
    4. two float64 shifts per pair are all-gathered into the reference's pair order.

    Returns ``(shift_y, shift_x)`` per pair (length n_seg (n_seg - 1) / 2) on every rank."""
    from . import _lib, imageprocess

    l = _lib.load()
    _rcc_decl(l)
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    dev = x.device
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    sb = segment_shards(n_seg, world)
    n_loc = sb[rank + 1] - sb[rank]
    n_max = max(sb[r + 1] - sb[r] for r in range(world))
    XH = X // 2 + 1
    Y_, X_, H, W = imageprocess._crop_geometry(Y, X, max_shift)
    seg_start = np.asarray(seg_start, dtype=np.int64)
    assert len(seg_start) == n_loc + 1
    p = lambda t: t.data_ptr() if t is not None and t.numel() else None     # noqa: E731
    marks = []

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))

    mark("start")
    # spectra of all ranks: rank r's segments at padded index r * n_max + k
    spectra = torch.empty((world * n_max, Y, XH, 2), dtype=torch.float32, device=dev)
    sums_all = torch.zeros(world * n_max, dtype=torch.float64, device=dev)
    if n_loc:
        segs = torch.empty((n_loc, Y, X), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        max_seg = int((seg_start[1:] - seg_start[:-1]).max()) if n_loc else 0
        wsb = int(l.pb_render_workspace_bytes(max_seg, Y, X))
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        for k in range(n_loc):
            a0, m = int(seg_start[k]), int(seg_start[k + 1] - seg_start[k])
            o = a0 * 4
            _lib.check(l.pb_render_dev(m, x.data_ptr() + o if m else None, y.data_ptr() + o if m else None,
                                       lpx.data_ptr() + o if m else None, lpy.data_ptr() + o if m else None,
                                       1.0, 0.0, 0.0, float(Y), float(X), float(min_blur_width), 1,
                                       segs[k].data_ptr(), Y, X, cnt.data_ptr(), ws.data_ptr(), wsb, st))
        mark("render")
        mine = spectra[rank * n_max: rank * n_max + n_loc]
        _lib.check(l.pb_rcc_spectra_dev(n_loc, Y, X, segs.data_ptr(), mine.data_ptr(),
                                        sums_all[rank * n_max:].data_ptr(), st))
        del segs, ws
    else:
        mark("render")
    mark("r2c")
    if world > 1:
        flat = spectra.view(-1)
        per = n_max * Y * XH * 2
        dist.all_gather_into_tensor(flat, flat[rank * per: (rank + 1) * per])          # in place
        s_mine = sums_all[rank * n_max: (rank + 1) * n_max].clone()
        dist.all_gather_into_tensor(sums_all, s_mine)
    mark("allgather_spectra")
    # pairs: whole L2 tiles per rank, indices mapped to the padded spectra layout
    TS = int(l.pb_rcc_tile_segments(Y))
    pi, pj, ref_index = my_tile_pairs(n_seg, TS, rank, world)
    seg_rank = np.searchsorted(np.asarray(sb[1:]), np.arange(n_seg), side="right")
    padded = (seg_rank * n_max + (np.arange(n_seg) - np.asarray(sb)[seg_rank])).astype(np.int32)
    n_mine = len(pi)
    rec = np.zeros((n_mine, 32), dtype=np.float64)
    win = None
    if n_mine:
        dpi = torch.as_tensor(padded[pi], device=dev)
        dpj = torch.as_tensor(padded[pj], device=dev)
        win = torch.empty((n_mine, H, W), dtype=torch.float32, device=dev)
        t_per_pair = H * XH * 8
        wsb = max(min(n_mine, 32768) * t_per_pair, Y * XH * 8 + Y * X * 4) + 4096
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(l.pb_rcc_windows_dev(n_mine, dpi.data_ptr(), dpj.data_ptr(), Y, X, spectra.data_ptr(),
                                        Y_, X_, H, W, win.data_ptr(), 1, ws.data_ptr(), wsb, st))
        mark("pairs")
        drec = torch.empty((n_mine, 32), dtype=torch.float64, device=dev)
        _lib.check(l.pb_rcc_peakfit_dev(n_mine, win.data_ptr(), H, W, drec.data_ptr(), st))
        rec = drec.cpu().numpy()
        del ws
    else:
        mark("pairs")
    mark("peakfit")
    sums = sums_all.cpu().numpy()[padded]                       # per real segment
    sy, sx = imageprocess._shifts_from_records(rec, sums, pi, pj, Y, X, Y_, X_,
                                               (lambda odd: win[torch.as_tensor(odd, device=dev)].cpu().numpy())
                                               if win is not None else None)
    n_pairs = n_seg * (n_seg - 1) // 2
    full = np.zeros((n_pairs, 2))
    if world > 1:
        b = shard_bounds(n_pairs, world)
        cmax = max(max(b[r + 1] - b[r] for r in range(world)), 1)
        pad = torch.zeros((cmax, 3), dtype=torch.float64, device=dev)
        if n_mine:
            pad[:n_mine] = torch.as_tensor(np.stack([sy, sx, ref_index.astype(np.float64)], 1), device=dev)
        out = torch.empty((world, cmax, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
        out = out.cpu().numpy()
        for r in range(world):
            c = b[r + 1] - b[r]
            full[out[r, :c, 2].astype(np.int64)] = out[r, :c, :2]
    else:
        full[ref_index] = np.stack([sy, sx], 1)
    mark("gather_shifts")
    if timings is not None:
        torch.cuda.synchronize(dev)
        for (na, a), (nb, b_) in zip(marks[:-1], marks[1:]):
            timings[nb + "_ms"] = a.elapsed_time(b_)
    return full[:, 0], full[:, 1]


def undrift_segments_sharded(dist, torch, locs, info, segmentation, device="cuda", timings=None):
    """``postprocess.undrift`` for a table sharded by FRAME RANGE over the ranks: ``locs`` (host
    DataFrame, frames ascending) holds the localizations of this rank's segments
    (``segment_shards`` of ``postprocess.n_segments`` segments; ``shard_locs_by_segment`` cuts a full
    table accordingly).  Device part: ``undrift_shifts_device``; then every rank runs the reference's
    ``minimize_shifts`` + cubic spline (milliseconds) and subtracts the drift from ITS rows only.
    Returns ``(drift DataFrame over all frames, undrifted shard)``."""
    import pandas as pd
    from scipy import interpolate

    from . import lib, postprocess

    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    n_frames = info[0]["Frames"]
    Y, X = info[0]["Height"], info[0]["Width"]
    n_seg = postprocess.n_segments(info, segmentation)
    bounds = np.linspace(0, n_frames - 1, n_seg + 1, dtype=np.uint32)
    sb = segment_shards(n_seg, world)
    frames = locs["frame"].to_numpy()
    if len(frames) > 1 and not bool(np.all(frames[1:] >= frames[:-1])):
        locs = locs.sort_values("frame", kind="stable")
        frames = locs["frame"].to_numpy()
    my_bounds = bounds[sb[rank]: sb[rank + 1] + 1]
    cut = np.searchsorted(frames, my_bounds, side="left")
    a, b = (int(cut[0]), int(cut[-1])) if len(cut) else (0, 0)
    seg_start = (cut - cut[0]).astype(np.int64) if len(cut) else np.zeros(1, np.int64)
    up = lambda c: _to_device(torch, np.ascontiguousarray(locs[c].to_numpy()[a:b], dtype=np.float32), device)   # noqa: E731
    x, y, lpx, lpy = up("x"), up("y"), up("lpx"), up("lpy")
    sy, sx = undrift_shifts_device(dist, torch, seg_start, x, y, lpx, lpy, n_seg, Y, X, min_blur_width=1.0,
                                   max_shift=32, timings=timings)
    shifts_x = np.zeros((n_seg, n_seg)); shifts_y = np.zeros((n_seg, n_seg))
    ai, aj = np.triu_indices(n_seg, 1)
    shifts_y[ai, aj] = sy; shifts_x[ai, aj] = sx
    shift_y, shift_x = lib.minimize_shifts(shifts_x, shifts_y)
    t = (bounds[1:] + bounds[:-1]) / 2
    t_inter = np.arange(n_frames)
    drift = pd.DataFrame({"x": interpolate.InterpolatedUnivariateSpline(t, shift_x, k=3)(t_inter),
                          "y": interpolate.InterpolatedUnivariateSpline(t, shift_y, k=3)(t_inter)})
    out = locs.copy()
    out = postprocess.apply_drift(out, info, drift=drift)
    return drift, out


def shard_locs_by_segment(locs, info, segmentation, rank, world):
    """The rows of a (frame-sorted) table that belong to this rank's segments; the last rank also
    keeps the frames past the final segment bound (the reference drops them from the images but
    still undrifts them)."""
    from . import postprocess

    n_frames = info[0]["Frames"]
    n_seg = postprocess.n_segments(info, segmentation)
    bounds = np.linspace(0, n_frames - 1, n_seg + 1, dtype=np.uint32)
    sb = segment_shards(n_seg, world)
    frames = locs["frame"].to_numpy()
    lo = np.searchsorted(frames, bounds[sb[rank]], side="left") if rank else 0
    hi = np.searchsorted(frames, bounds[sb[rank + 1]], side="left") if rank + 1 < world else len(frames)
    return locs.iloc[int(lo): int(hi)]


# ---- fused localize on a device-resident frame block ---------------------------------------------
def localize_device(torch, movie, frame_offset, camera_info, parameters, *, fitting_method="gausslq",
                    eps=0.001, max_it=100, mle_method="sigmaxy", roi=None):
    """``localize.localize`` on a frame block that already lives in HBM (``movie``: (F, Y, X) uint16
    or float32 CUDA tensor; ``frame_offset`` = index of its first frame in the whole movie):
    identify -> sort -> cut -> fit -> localization columns without leaving the GPU
    (``pb_localize_dev``).  Returns the (ncols, n) float32 column block as a CUDA tensor, rows in
    (frame, y, x) order; ``localize._columns_to_locs`` turns a downloaded block into the DataFrame."""
    import ctypes as C

    from . import _lib, localize as pbl

    l = _lib.load()
    vp, i32, sz, f32, f64 = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_double
    l.pb_localize_dev.argtypes = [vp, i32, sz, i32, i32, C.c_longlong, i32, f64, vp, f32, f32, f32, i32, f64, i32,
                                  i32, vp, sz, C.POINTER(sz)]
    l.pb_localize_dev.restype = i32
    fit = pbl._FIT_IDS[(fitting_method, mle_method if fitting_method == "gaussmle" else None)]
    ncols = int(l.pb_locs_columns(fit))
    F, Y, X = movie.shape
    if movie.dtype in (torch.int16, getattr(torch, "uint16", torch.int16)):
        dtype = 0                      # uint16 counts (int16 storage is read as uint16)
    elif movie.dtype == torch.float32:
        dtype = 1
    else:
        raise ValueError("movie must be a uint16 / int16 or float32 CUDA tensor")
    roi_arr = pbl._roi_array(roi)
    capacity = max(4096, 128 * int(F))
    while True:
        cols = torch.empty((ncols, capacity), dtype=torch.float32, device=movie.device)
        found = sz(0)
        rc = l.pb_localize_dev(movie.data_ptr(), dtype, int(F), int(Y), int(X), int(frame_offset),
                               int(parameters["Box Size"]), float(parameters["Min. Net Gradient"]),
                               _lib.ptr(roi_arr) if roi_arr is not None else None,
                               float(camera_info["Baseline"]), float(camera_info["Sensitivity"]),
                               float(camera_info["Gain"]), fit, float(eps), int(max_it),
                               int(camera_info["Gain"] > 1), cols.data_ptr(), capacity, C.byref(found))
        if rc == 4:
            capacity = int(found.value)
            continue
        _lib.check(rc)
        return cols[:, : int(found.value)]


# ---- NVSwitch multicast buffer (NVLS): the fused fit + all-gather ---------------------------------
class MulticastBuffer:
    """A gather buffer of ``world`` equal blocks that exists on every rank and is written THROUGH THE
    SWITCH: memory of all GPUs is bound to one NVSwitch multicast object (csrc/multicast.cu), so a store
    to ``mc_ptr + offset`` lands at ``uc_ptr + offset`` on every rank.  ``pb_mle_fit_gather_dev`` stores
    each finished spot's results through ``block_mc_ptr()`` -- the all-gather of the fit results is part
    of the fit kernel (one store stream per rank, 1/(N-1) of the NVLink egress of unicast copies, no copy
    engine, no collective kernel).

    One process per GPU: rank 0 creates the object and hands its POSIX file descriptor to the other
    ranks over a Unix-domain socket (SCM_RIGHTS); ``torch.distributed`` carries only the socket path and
    the barriers.  Raises ``RuntimeError`` when the GPUs / driver have no multicast support."""

    def __init__(self, dist, torch, block_bytes: int, device):
        import ctypes as C
        import os
        import socket
        import tempfile

        from . import _lib

        self.dist, self.torch, self._lib = dist, torch, _lib
        self.l = l = _lib.load()
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        l.pb_mc_padded_size.argtypes = [sz, i32, C.POINTER(sz)]
        l.pb_mc_create.argtypes = [sz, i32, C.POINTER(vp), C.POINTER(i32)]
        l.pb_mc_import.argtypes = [i32, sz, i32, C.POINTER(vp)]
        l.pb_mc_add_device.argtypes = [vp]
        l.pb_mc_bind_map.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        l.pb_mc_destroy.argtypes = [vp]
        l.pb_mc_copy_async.argtypes = [vp, vp, sz, i32, vp]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.block_bytes = int(block_bytes)
        self.device = device
        self.handle = None
        ok = torch.tensor([float(l.pb_mc_supported())], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 0.5:
            raise RuntimeError("NVSwitch multicast (NVLS) is not supported on every rank")
        padded = sz(0)
        _lib.check(l.pb_mc_padded_size(self.block_bytes * self.world, self.world, C.byref(padded)))
        self.nbytes = int(padded.value)
        h = vp()
        status = torch.ones(1, device=device)
        path = [None]
        srv = None
        try:
            if self.rank == 0:
                fd = i32(-1)
                _lib.check(l.pb_mc_create(self.nbytes, self.world, C.byref(h), C.byref(fd)))
                path[0] = os.path.join(tempfile.gettempdir(), f"pb_mc_{os.getpid()}_{id(self) & 0xffffff:x}.sock")
                if os.path.exists(path[0]):
                    os.unlink(path[0])
                srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                srv.bind(path[0])
                srv.listen(self.world)
            dist.broadcast_object_list(path, src=0)
            if self.rank == 0:
                for _ in range(self.world - 1):
                    conn, _a = srv.accept()
                    socket.send_fds(conn, [b"mc"], [fd.value])
                    conn.recv(2)                      # the peer has imported the handle
                    conn.close()
                os.close(fd.value)
            else:
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                c.connect(path[0])
                _msg, fds, _f, _a = socket.recv_fds(c, 16, 1)
                _lib.check(l.pb_mc_import(fds[0], self.nbytes, self.world, C.byref(h)))
                c.send(b"ok")
                c.close()
        except Exception:
            status.zero_()
            raise
        finally:
            if srv is not None:
                srv.close()
                try:
                    os.unlink(path[0])
                except OSError:
                    pass
        self.handle = h
        _lib.check(l.pb_mc_add_device(h))
        dist.barrier()                                # every device has joined the team
        uc, mc = vp(), vp()
        _lib.check(l.pb_mc_bind_map(h, C.byref(uc), C.byref(mc)))
        self.uc_ptr, self.mc_ptr = uc.value, mc.value
        dist.barrier()                                # every rank's memory is bound before anyone stores

    def block_mc_ptr(self, rank=None) -> int:
        """Multicast address of a rank's block (default: this rank's)."""
        return self.mc_ptr + (self.rank if rank is None else rank) * self.block_bytes

    def local(self, dtype):
        """This rank's copy of the whole gather buffer as a torch tensor (no copy)."""
        torch = self.torch
        itemsize = torch.empty(0, dtype=dtype).element_size()
        n = self.block_bytes * self.world // itemsize

        class _Raw:
            pass

        raw = _Raw()
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.int64: "<i8"}[dtype]
        raw.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (self.uc_ptr, False), "version": 2}
        t = torch.as_tensor(raw, device=self.device)
        self._keep = raw
        return t

    def copy_in(self, block, stream=None, n_ctas: int = 16):
        """Push an existing device tensor (this rank's block) through the multicast mapping."""
        torch = self.torch
        st = stream or torch.cuda.current_stream(self.device)
        nbytes = block.numel() * block.element_size()
        assert block.is_contiguous() and nbytes <= self.block_bytes and nbytes % 16 == 0
        self._lib.check(self.l.pb_mc_copy_async(self.block_mc_ptr(), block.data_ptr(), nbytes, n_ctas, st.cuda_stream))

    def close(self):
        if self.handle is not None:
            self.torch.cuda.synchronize(self.device)
            self.dist.barrier()
            self.l.pb_mc_destroy(self.handle)
            self.handle = None
