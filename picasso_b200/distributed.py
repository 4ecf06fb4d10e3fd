"""Multi-GPU plumbing over ``torch.distributed`` (one process per GPU; NCCL on B200,
gloo in the CPU tests).  SURVEY.md section 8e: every stage of the hot path shards without
a data-path collective; the only exchanges are the gather of results.

* fits: spots shard by contiguous index range; ONE all-gather of the packed per-rank
  output buffer ``[thetas 6n | crlbs 6n | logliks n | iterations n]``.
* identify: frames shard by contiguous range; variable-length results are exchanged by an
  all-gather of counts followed by a padded all-gather.
* render: localisations shard by index, the partial images are summed (all-reduce).
* fused localize: frames shard by contiguous block, the finished column blocks are gathered.
* undrift: every rank renders and transforms all segments (33 ms for 200 x 4096^2), the
  n(n-1)/2 pairs shard round-robin; only the 4 KB correlation windows are exchanged.
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

    def gather_async(self, block, stream=None, offset_bytes: int = 0):
        """Enqueue the copies of ``block`` (a contiguous device tensor; the whole block of this rank,
        or the part of it that starts ``offset_bytes`` into the block) into every rank's buffer,
        ordered after the work already enqueued on ``stream``; returns the events that complete
        when the data has left this rank."""
        torch = self.torch
        stream = stream or torch.cuda.current_stream(self.device)
        nbytes = block.numel() * block.element_size()
        assert block.is_contiguous() and offset_bytes + nbytes <= self.block_bytes
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
