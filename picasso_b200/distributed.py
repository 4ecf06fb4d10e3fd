"""Multi-GPU plumbing over ``torch.distributed`` (one process per GPU; NCCL on B200,
gloo in the CPU tests).  SURVEY.md section 8e: every stage of the hot path shards without
a data-path collective; the only exchanges are the gather of results.

* fits: spots shard by contiguous index range; ONE all-gather of the packed per-rank
  output buffer ``[thetas 6n | crlbs 6n | logliks n | iterations n]``.
* identify: frames shard by contiguous range; variable-length results are exchanged by an
  all-gather of counts followed by a padded all-gather.
* render: localisations shard by index, the partial images are summed (all-reduce).
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
