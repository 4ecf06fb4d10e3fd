"""Multi-GPU check + timing of the sharded stages under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_multi.py [--small]

Every stage is run sharded over all ranks (NCCL) and, on rank 0, on one GPU; the results must
agree (tables / drift bit for bit, images to summation order).  One JSON line on rank 0."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
warnings.filterwarnings("ignore")


def main():
    import pandas as pd
    import torch
    import torch.distributed as dist

    from bench_stages import gen_movie_device
    from picasso_b200 import _lib, distributed as pbd, localize, postprocess, render, testing

    small = "--small" in sys.argv
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    _lib.check(_lib.load().pb_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def timed(fn, reps=3):
        """Best of `reps` (the first full-size call grows the cached pinned / device buffers and
        sets up NCCL channels for the message size); max over ranks of each repeat."""
        best = None
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize(); dist.barrier()
            t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return r, best

    def timed_one(fn, reps=3):
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return r, best

    out = {"world": world}
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    par = {"Min. Net Gradient": 5000, "Box Size": 7}
    # ---- fused localize, frames sharded ----------------------------------------------------
    F = 400 if small else 2000
    movie = gen_movie_device(torch, F, 512, 512, dev=dev).cpu().numpy()
    pbd.localize_sharded(dist, torch, movie[:16], dict(cam), par, fitting_method="gaussmle", device=dev)
    locs, t_sh = timed(lambda: pbd.localize_sharded(dist, torch, movie, dict(cam), par,
                                                    fitting_method="gaussmle", device=dev))
    res = {"frames": F, "sharded_seconds": t_sh, "n_locs": len(locs)}
    if rank == 0:
        one, res["one_gpu_seconds"] = timed_one(lambda: localize.localize(
            movie, dict(cam), par, fitting_method="gaussmle", return_info=False))
        res["tables_bit_identical"] = all(locs[c].to_numpy().tobytes() == one[c].to_numpy().tobytes()
                                          for c in one.columns) and len(one) == len(locs)
    out["localize"] = res
    del movie
    # ---- render, localizations sharded + image all-reduce -----------------------------------
    n = 4_000_000 if small else 20_000_000
    rng = np.random.default_rng(2)
    rl = pd.DataFrame({"x": rng.uniform(0, 512, n).astype(np.float32), "y": rng.uniform(0, 512, n).astype(np.float32),
                       "lpx": rng.uniform(0.02, 0.08, n).astype(np.float32),
                       "lpy": rng.uniform(0.02, 0.08, n).astype(np.float32)})
    info = [{"Height": 512, "Width": 512, "Frames": 1, "Pixelsize": 130}]
    kw = dict(oversampling=20, blur_method="gaussian")
    pbd.render_sharded(dist, torch, rl.iloc[:1000], info, device=dev, **kw)
    (k, img), t_sh = timed(lambda: pbd.render_sharded(dist, torch, rl, info, device=dev, **kw))
    res = {"n_locs": n, "sharded_seconds": t_sh, "n_in_view": k}
    if rank == 0:
        (k1, img1), res["one_gpu_seconds"] = timed_one(lambda: render.render(rl, info, **kw))
        big = img1 > 1e-3 * img1.max()
        res["n_equal"] = bool(k1 == k)
        res["max_rel_dev_bright_pixels"] = float(np.max(np.abs(img[big] - img1[big]) / img1[big]))
    out["render"] = res
    del rl, img
    # ---- undrift, pairs sharded ----------------------------------------------------------------
    nf, side = (4000, 1024) if small else (12000, 2048)
    dl, dinfo, truth = testing.synthetic_drift_locs(nf, side, side, n_clusters=2000, locs_per_frame=500.0, seed=3)
    pbd.undrift_sharded(dist, torch, dl[dl["frame"] < 500], [{**dinfo[0], "Frames": 500}], 100, device=dev)
    (drift, und), t_sh = timed(lambda: pbd.undrift_sharded(dist, torch, dl, dinfo, 100, device=dev))
    res = {"n_locs": len(dl), "segments": nf // 100, "image": [side, side], "sharded_seconds": t_sh}
    if rank == 0:
        (d1, u1), res["one_gpu_seconds"] = timed_one(lambda: postprocess.undrift(
            dl, dinfo, 100, display=False, segmentation_callback=lambda i: None, rcc_callback=lambda i: None))
        res["drift_bit_identical"] = bool(d1["x"].to_numpy().tobytes() == drift["x"].to_numpy().tobytes()
                                          and d1["y"].to_numpy().tobytes() == drift["y"].to_numpy().tobytes())
        res["max_abs_drift_dev"] = float(max(np.abs(d1["x"] - drift["x"]).max(), np.abs(d1["y"] - drift["y"]).max()))
    out["undrift"] = res
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
