"""A/B the MLE kernel families on the GPU box (pb_mle_set_impl): timing on device-resident
config-2 spots and parity on 100 k spots against the CPU oracle.

    python tools/bench_mle_impls.py [n_spots] [impls...]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from picasso_b200 import _lib, testing  # noqa: E402

NAMES = {0: "lane-group (mle_fit.cu)", 1: "thread-per-spot f64 pixels", 2: "thread-per-spot f32 pixels"}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    impls = [int(a) for a in sys.argv[2:]] or [0, 1, 2]
    lib = _lib.load()
    _lib.require_gpu()
    dev = torch.device("cuda", 0)
    spots = bench.gen_spots_device(torch, n, 7, 1234, dev)
    par = testing.synthetic_spots(100000, 7, seed=3)
    oth, ocr, oll, oit = oracle.gaussmle(par, 0.001, 100, "sigmaxy", nthreads=os.cpu_count())
    dpar = torch.from_numpy(par).to(dev)
    th = torch.empty((n, 6), device=dev)
    cr = torch.empty((n, 6), device=dev)
    ll = torch.empty(n, device=dev)
    it = torch.empty(n, dtype=torch.int32, device=dev)

    def go(sp, m):
        _lib.check(lib.pb_mle_fit_dev(m, 7, sp.data_ptr(), 0.001, 100, 1, th.data_ptr(), cr.data_ptr(),
                                      ll.data_ptr(), it.data_ptr(), None, None))

    for impl in impls:
        _lib.check(lib.pb_mle_set_impl(impl))
        for _ in range(3):
            go(spots, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(5):
            go(spots, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        launches = (_lib.launch_count() - l0) // 5
        mean_it = float(it.float().mean().item())
        go(dpar, len(par))
        torch.cuda.synchronize()
        m = len(par)
        pit = it[:m].cpu().numpy()
        pth = th[:m].cpu().numpy()
        pcr = cr[:m].cpu().numpy()
        pll = ll[:m].cpu().numpy()
        same = pit == oit
        d = pth.astype(np.float64) - oth
        rms = np.sqrt((d ** 2).mean(0))
        with np.errstate(divide="ignore", invalid="ignore"):
            crl = np.abs(pcr - ocr) / np.abs(ocr)
        print(json.dumps({
            "impl": impl, "name": NAMES[impl], "n": n, "ms": ms, "Mfits_per_s": n / ms / 1e3,
            "launches_per_call": launches, "mean_iterations": mean_it,
            "iter_match": float(same.mean()), "rms_x": rms[0], "rms_y": rms[1], "rms_sx": rms[4],
            "rms_sy": rms[5], "rel_rms_photons": float(np.sqrt(((d[:, 2] / oth[:, 2]) ** 2).mean())),
            "rel_rms_bg": float(np.sqrt(((d[:, 3] / oth[:, 3]) ** 2).mean())),
            "theta_bit_identical": float((pth.view(np.uint32) == oth.view(np.uint32)).all(1).mean()),
            "max_abs_xysigma_same_iter": float(np.abs(d[same][:, [0, 1, 4, 5]]).max()),
            "crlb_rel_median": float(np.nanmedian(crl[same])),
            "loglik_max_abs_same_iter": float(np.abs(pll[same] - oll[same]).max()),
        }), flush=True)


if __name__ == "__main__":
    main()
