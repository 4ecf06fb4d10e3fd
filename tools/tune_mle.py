"""Build MLE kernel variants (-D knobs) into picasso_b200/_variants/ and time them.

    python tools/tune_mle.py build      # here (no GPU): compiles the variants
    python tools/tune_mle.py run        # on the GPU box: times each variant
"""
import ctypes as C
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "picasso_b200", "_variants")
CSRC = os.path.join(ROOT, "picasso_b200", "csrc")
NVCC = "/usr/local/cuda/bin/nvcc"

VARIANTS = {
    "default": [],
    "b5": ["-DPB_MLE_MINB=5"],
    "f32pixels": ["-DPB_MLE_F32_PIXELS=1"],
    "f32pixels_b5": ["-DPB_MLE_F32_PIXELS=1", "-DPB_MLE_MINB=5"],
}


def build():
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", "-Xptxas", "-v", *flags,
               os.path.join(CSRC, "mle_fit.cu"), os.path.join(CSRC, "api.cu"), "-o", out]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        outp = p.communicate()[0]
        if p.returncode:
            print(name, "FAILED\n", outp[-2000:])
            continue
        # report registers / spills of the box-7 sigmaxy kernel
        lines = outp.splitlines()
        for i, ln in enumerate(lines):
            if "mle_fit_kernelILi7ELi8ELi1E" in ln and "Compiling" in ln:
                print(name, lines[i + 2].strip(), "|", lines[i + 3].strip())


def run():
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    import bench
    import oracle
    from picasso_b200 import testing

    n = 2_000_000
    dev = torch.device("cuda", 0)
    spots = bench.gen_spots_device(torch, n, 7, 1234, dev)
    par = testing.synthetic_spots(100000, 7, seed=3)
    oth, ocr, oll, oit = oracle.gaussmle(par, 0.001, 100, "sigmaxy", nthreads=os.cpu_count())
    dpar = torch.from_numpy(par).to(dev)
    for name in VARIANTS:
        path = os.path.join(VDIR, f"lib_{name}.so")
        if not os.path.exists(path):
            continue
        lib = C.CDLL(path)
        vp = C.c_void_p
        lib.pb_mle_fit_dev.argtypes = [C.c_size_t, C.c_int, vp, C.c_double, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
        th = torch.empty((n, 6), device=dev); cr = torch.empty((n, 6), device=dev)
        ll = torch.empty(n, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)

        def go(sp, m):
            rc = lib.pb_mle_fit_dev(m, 7, sp.data_ptr(), 0.001, 100, 1, th.data_ptr(), cr.data_ptr(),
                                    ll.data_ptr(), it.data_ptr(), None, None)
            assert rc == 0
        for _ in range(3):
            go(spots, n)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            go(spots, n)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        go(dpar, len(par)); torch.cuda.synchronize()
        pit = it[:len(par)].cpu().numpy(); pth = th[:len(par)].cpu().numpy()
        match = float((pit == oit).mean())
        rms = float(np.sqrt(((pth[:, :2] - oth[:, :2]).astype(np.float64) ** 2).mean()))
        bit = float((pth.view(np.uint32) == oth.view(np.uint32)).all(1).mean())
        print(json.dumps({"variant": name, "ms": ms, "Mfits_per_s": n / ms / 1e3,
                          "iter_match": match, "xy_rms": rms, "theta_bit_identical": bit}), flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
