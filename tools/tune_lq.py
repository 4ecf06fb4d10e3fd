import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from picasso_b200 import testing
n = 400_000
spots = torch.from_numpy(np.tile(testing.synthetic_spots(20000, 7, seed=5), (20, 1, 1))).cuda()
for t in ("ne_t64", "ne_t128", "ne_t256", "qr_t128"):
    lib = C.CDLL(os.path.join(ROOT, "picasso_b200", "_variants", f"lib_lq_{t}.so"))
    vp = C.c_void_p
    lib.pb_lq_fit_dev.argtypes = [C.c_size_t, C.c_int, vp, vp, vp, vp, vp]
    th = torch.empty((n, 6), device="cuda")
    go = lambda: lib.pb_lq_fit_dev(n, 7, spots.data_ptr(), th.data_ptr(), None, None, None)
    for _ in range(2): go()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"variant": t, "ms": ms, "Mfits_per_s": n / ms / 1e3}), flush=True)
