"""N-GPU check (torchrun) of the scalable sharded stages of picasso_b200.distributed against the
single-GPU product calls on the same data:

  render_bands (row bands + all-to-all)      vs render.render                  n equal, pixels 1e-4 rel
  undrift_segments_sharded (segments + spectra all-gather + tile-sharded pairs)
                                             vs postprocess.undrift            drift 1e-5 px, rows equal

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tools/check_sharded_scalable.py

One JSON line on rank 0."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def main():
    import pandas as pd
    import torch
    import torch.distributed as dist

    from picasso_b200 import _lib, distributed as pbd, postprocess, render, testing

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    _lib.check(_lib.load().pb_set_device(local))
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    # ---- render by row bands ------------------------------------------------------------------
    rng = np.random.default_rng(2)
    n = 1_500_000
    rl = pd.DataFrame({"x": rng.uniform(-2, 130, n).astype(np.float32), "y": rng.uniform(-2, 130, n).astype(np.float32),
                       "lpx": rng.uniform(0.02, 0.3, n).astype(np.float32),
                       "lpy": rng.uniform(0.02, 0.3, n).astype(np.float32)})
    info = [{"Height": 128, "Width": 128, "Frames": 1, "Pixelsize": 130}]
    lo, hi = pbd.my_shard(n, rank, world)
    res = {}
    for bm in ("gaussian", "gaussian_iso", None):
        k, band, (r0, r1) = pbd.render_bands(dist, torch, rl.iloc[lo:hi], info, device=dev, oversampling=20,
                                             blur_method=bm)
        k1, img1 = render.render(rl, info, oversampling=20, blur_method=bm)
        ref = img1[r0:r1]
        if bm is None:
            ok = bool(np.array_equal(band, ref))
            rel = 0.0
        else:
            big = ref > 1e-3 * img1.max()
            rel = float((np.abs(band - ref)[big] / ref[big]).max()) if big.any() else 0.0
            ok = rel <= 1e-4
        t = torch.tensor([1.0 if (ok and k == k1) else 0.0, rel], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if False else dist.ReduceOp.MAX)
        flag = torch.tensor([1.0 if (ok and k == k1) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        full = pbd.gather_bands(dist, torch, band, img1.shape[0], device=dev)
        res[str(bm)] = {"ok": bool(flag.item() > 0.5), "max_rel": float(t[1].item()), "n": int(k), "n_ref": int(k1),
                        "gathered_shape_ok": bool(full.shape == img1.shape),
                        "gathered_sum_rel": float(abs(full.sum(dtype=np.float64) - img1.sum(dtype=np.float64))
                                                  / max(img1.sum(dtype=np.float64), 1e-30))}
    out["render_bands"] = res

    # ---- undrift: segments sharded ----------------------------------------------------------
    nf, side = 3000, 1024
    dl, dinfo, truth = testing.synthetic_drift_locs(nf, side, side, n_clusters=600, locs_per_frame=60.0, seed=3)
    mine = pbd.shard_locs_by_segment(dl, dinfo, 100, rank, world)
    drift, und = pbd.undrift_segments_sharded(dist, torch, mine, dinfo, 100, device=dev)
    d1, u1 = postprocess.undrift(dl, dinfo, 100, display=False, segmentation_callback=lambda i: None,
                                 rcc_callback=lambda i: None)
    dev_drift = float(max(np.abs(d1["x"] - drift["x"]).max(), np.abs(d1["y"] - drift["y"]).max()))
    ref_rows = u1.loc[mine.index]
    dev_rows = float(max(np.abs(ref_rows["x"].to_numpy() - und["x"].to_numpy()).max(),
                         np.abs(ref_rows["y"].to_numpy() - und["y"].to_numpy()).max())) if len(mine) else 0.0
    t = torch.tensor([dev_drift, dev_rows, float(len(und))], device=dev, dtype=torch.float64)
    tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    out["undrift_segments"] = {"max_abs_drift_dev_px": float(tm[0].item()), "max_abs_row_dev_px": float(tm[1].item()),
                               "rows_total": int(ts[2].item()), "rows_ref": int(len(u1)),
                               "drift_vs_injected_px": float(np.abs((d1["x"] - d1["x"].mean())
                                                                    - (truth[:, 0] - truth[:, 0].mean())).max())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
