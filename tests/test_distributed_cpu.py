"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: index sharding, the single
all-gather of packed fit outputs, variable-length gathers and the image all-reduce.
The per-rank compute is the CPU oracle here (test infrastructure); on B200 it is the CUDA
library -- the plumbing is backend-agnostic."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from picasso_b200 import distributed as pbd
    from picasso_b200 import testing

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spots = testing.synthetic_spots(1001, 7, seed=4)     # odd count: uneven shards
        fit = lambda s: oracle.gaussmle(s, 0.001, 100, "sigmaxy")
        th, cr, ll, it = pbd.sharded_fit(dist, torch, spots, fit)
        # variable-length gather: each rank identifies its share of frames
        movie = testing.synthetic_movie(6, 48, 48, emitters_per_frame=4, seed=9)
        lo, hi = pbd.my_shard(len(movie), rank, world)
        fr, x, y, ng = oracle.identify_movie(movie[lo:hi], 3000, 7)
        fr = fr + lo
        g = pbd.all_gather_variable(dist, torch, (fr, x, y, ng))
        # image all-reduce: render shards of the localisations
        rng = np.random.default_rng(1)
        locs = {"x": rng.uniform(0, 16, 500).astype(np.float32),
                "y": rng.uniform(0, 16, 500).astype(np.float32),
                "lpx": np.full(500, 0.1, np.float32), "lpy": np.full(500, 0.1, np.float32)}
        info = [{"Height": 16, "Width": 16, "Pixelsize": 100}]
        lo, hi = pbd.my_shard(500, rank, world)
        part = {k: v[lo:hi] for k, v in locs.items()}
        n_part, img = oracle.render(part, info, oversampling=4, blur_method="gaussian")
        img = pbd.all_reduce_image(dist, torch, img)
        # column blocks of the fused localize path (uint32 bit patterns in a float32 block)
        n_r = 5 + 3 * rank
        blk = np.arange(3 * n_r, dtype=np.float32).reshape(3, n_r) + 100 * rank
        blk[0] = (np.arange(n_r, dtype=np.uint32) + 0x7FC00000 + rank).view(np.float32)   # NaN payloads
        cols = pbd.gather_column_blocks(dist, torch, blk)
        # pair windows sharded round-robin (undrift)
        n_seg = 7
        pi, pj = pbd.my_pairs(n_seg, rank, world)
        win = np.stack([np.full((4, 4), 100 * i + j, np.float32) for i, j in zip(pi, pj)])
        full = pbd.gather_pair_windows(dist, torch, win, n_seg)
        vals = pbd.gather_pair_values(dist, torch, np.stack([pi * 1.5, pj * 2.5], 1), n_seg)
        assert np.array_equal(vals[:, 0], np.triu_indices(n_seg, 1)[0] * 1.5)
        assert np.array_equal(vals[:, 1], np.triu_indices(n_seg, 1)[1] * 2.5)
        if rank == 0:
            q.put((th, cr, ll, it, g, img, cols, full))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing(oracle):
    import torch.multiprocessing as mp

    from picasso_b200 import testing

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    th, cr, ll, it, g, img, cols, full = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spots = testing.synthetic_spots(1001, 7, seed=4)
    oth, ocr, oll, oit = oracle.gaussmle(spots, 0.001, 100, "sigmaxy")
    np.testing.assert_array_equal(th, oth)
    np.testing.assert_array_equal(cr, ocr)
    np.testing.assert_array_equal(ll, oll)
    np.testing.assert_array_equal(it, oit)
    movie = testing.synthetic_movie(6, 48, 48, emitters_per_frame=4, seed=9)
    fr, x, y, ng = oracle.identify_movie(movie, 3000, 7)
    np.testing.assert_array_equal(g[0], fr)
    np.testing.assert_array_equal(g[1], x)
    np.testing.assert_array_equal(g[3], ng)
    rng = np.random.default_rng(1)
    locs = {"x": rng.uniform(0, 16, 500).astype(np.float32),
            "y": rng.uniform(0, 16, 500).astype(np.float32),
            "lpx": np.full(500, 0.1, np.float32), "lpy": np.full(500, 0.1, np.float32)}
    n_all, ref = oracle.render(locs, [{"Height": 16, "Width": 16, "Pixelsize": 100}],
                               oversampling=4, blur_method="gaussian")
    np.testing.assert_allclose(img, ref, rtol=1e-5, atol=1e-7)
    # gathered column blocks: rank order, payload bits preserved
    assert cols.shape == (3, 5 + 8)
    np.testing.assert_array_equal(cols[0].view(np.uint32)[:5], np.arange(5, dtype=np.uint32) + 0x7FC00000)
    np.testing.assert_array_equal(cols[0].view(np.uint32)[5:], np.arange(8, dtype=np.uint32) + 0x7FC00001)
    np.testing.assert_array_equal(cols[1][5:], np.arange(8, 16, dtype=np.float32) + 100)
    # gathered pair windows: the reference's pair order (i outer, j inner)
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 7)]
    assert full.shape == (21, 4, 4)
    np.testing.assert_array_equal(full[:, 0, 0], [100 * i + j for i, j in pairs])


def test_shard_bounds():
    from picasso_b200 import distributed as pbd

    assert pbd.shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    assert pbd.shard_bounds(0, 3) == [0, 0, 0, 0]
    for n in (1, 7, 1000, 10_000_001):
        b = pbd.shard_bounds(n, 8)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(8))
        assert max(b[i + 1] - b[i] for i in range(8)) - min(b[i + 1] - b[i] for i in range(8)) <= 1


# ---- row-band render exchange and the undrift partitions (scalable sharding) -------------------
def _band_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from picasso_b200 import distributed as pbd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        n, os_, H, W = 4000, 8.0, 40, 24
        cols = {"x": rng.uniform(-1, W + 1, n).astype(np.float32), "y": rng.uniform(-1, H + 1, n).astype(np.float32),
                "lpx": rng.uniform(0.05, 0.4, n).astype(np.float32), "lpy": rng.uniform(0.05, 0.4, n).astype(np.float32)}
        lo, hi = pbd.my_shard(n, rank, world)
        mine = {k: v[lo:hi] for k, v in cols.items()}
        npy = int(np.ceil(os_ * H))
        rows = pbd.band_rows(npy, world)
        # numpy mirror of the device bucketing (csrc/render.cu band_range): every band the 3 sigma + 2
        # row window reaches
        y_ = os_ * mine["y"].astype(np.float64)
        inview = (mine["x"] > 0) & (mine["y"] > 0) & (mine["x"] < W) & (mine["y"] < H)
        reach = 3.0 * (np.float32(os_) * np.maximum(mine["lpx"], mine["lpy"])).astype(np.float64) + 2.0
        send, counts = {k: [] for k in mine}, []
        for b in range(world):
            sel = inview & (y_ + reach >= rows[b]) & (y_ - reach <= rows[b + 1]) & (rows[b + 1] > rows[b])
            counts.append(int(sel.sum()))
            for k in mine:
                send[k].append(mine[k][sel])
        allc = pbd.all_gather_counts(dist, torch, counts, "cpu")
        recv = {k: pbd.exchange_variable(dist, torch, torch.from_numpy(np.concatenate(send[k])), counts, allc).numpy()
                for k in mine}
        info = [{"Height": H, "Width": W, "Pixelsize": 100}]
        _, img = oracle.render(recv, info, oversampling=os_, blur_method="gaussian")
        band = img[rows[rank]: rows[rank + 1]]
        full = pbd.gather_bands(dist, torch, band, npy)
        if rank == 0:
            q.put((full, allc))
    finally:
        dist.destroy_process_group()


def test_row_band_exchange_gloo(oracle):
    import torch.multiprocessing as mp

    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, allc = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    n, H, W = 4000, 40, 24
    cols = {"x": rng.uniform(-1, W + 1, n).astype(np.float32), "y": rng.uniform(-1, H + 1, n).astype(np.float32),
            "lpx": rng.uniform(0.05, 0.4, n).astype(np.float32), "lpy": rng.uniform(0.05, 0.4, n).astype(np.float32)}
    n_all, ref = oracle.render(cols, [{"Height": H, "Width": W, "Pixelsize": 100}], oversampling=8.0,
                               blur_method="gaussian")
    assert full.shape == ref.shape
    np.testing.assert_allclose(full, ref, rtol=1e-5, atol=1e-7)
    assert allc.shape == (3, 3) and allc.sum() >= n_all        # halo localizations are sent twice


def test_band_rows_and_pair_tiles():
    from picasso_b200 import distributed as pbd

    for npy, world in ((10240, 8), (10240, 1), (100, 4), (640, 3), (64, 8), (4097, 5)):
        r = pbd.band_rows(npy, world)
        assert len(r) == world + 1 and r[0] == 0 and r[-1] == npy
        assert all(r[i] <= r[i + 1] for i in range(world))
        assert all(v % 64 == 0 for v in r[1:-1])
    assert pbd.band_rows(10240, 8) == [1280 * k for k in range(9)]
    # pair tiles: every pair exactly once over the ranks, whole tiles contiguous
    for n_seg, ts, world in ((200, 20, 8), (20, 4, 3), (7, 2, 2), (5, 64, 4)):
        seen = np.concatenate([pbd.my_tile_pairs(n_seg, ts, r, world)[2] for r in range(world)])
        assert sorted(seen.tolist()) == list(range(n_seg * (n_seg - 1) // 2))
        pi, pj, order = pbd.tile_sorted_pairs(n_seg, ts)
        ai, aj = np.triu_indices(n_seg, 1)
        np.testing.assert_array_equal(ai[order], pi)
        np.testing.assert_array_equal(aj[order], pj)
        key = (pi // ts) * (n_seg + 1) + pj // ts
        assert (np.diff(key) >= 0).all()


def test_shard_locs_by_segment():
    import pandas as pd

    from picasso_b200 import distributed as pbd, postprocess

    info = [{"Height": 64, "Width": 64, "Frames": 1000, "Pixelsize": 130}]
    rng = np.random.default_rng(0)
    frames = np.sort(rng.integers(0, 1000, 5000)).astype(np.uint32)
    locs = pd.DataFrame({"frame": frames, "x": rng.uniform(0, 64, 5000).astype(np.float32)})
    n_seg = postprocess.n_segments(info, 100)
    bounds = np.linspace(0, 999, n_seg + 1, dtype=np.uint32)
    parts = [pbd.shard_locs_by_segment(locs, info, 100, r, 4) for r in range(4)]
    assert sum(len(p) for p in parts) == len(locs)                     # a partition of the rows
    pd.testing.assert_frame_equal(pd.concat(parts), locs)
    sb = pbd.segment_shards(n_seg, 4)
    for r, p in enumerate(parts):
        f = p["frame"].to_numpy()
        if r:
            assert f.min() >= bounds[sb[r]]
        if r + 1 < 4:
            assert f.max() < bounds[sb[r + 1]]
