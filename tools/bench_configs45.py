"""BASELINE.json configs 4 and 5 end to end through the Python API (one GPU):
render.render of 50 M localizations at oversampling 20, postprocess.undrift of 20 M localizations
in 20 000 frames of 4096 x 4096 (segmentation 100 -> 200 segments, 19 900 pairs)."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def main():
    import pandas as pd

    from picasso_b200 import postprocess, render, testing

    out = {}
    rng = np.random.default_rng(2)
    n = 50_000_000
    locs = pd.DataFrame({"x": rng.uniform(0, 512, n).astype(np.float32), "y": rng.uniform(0, 512, n).astype(np.float32),
                         "lpx": rng.uniform(0.02, 0.08, n).astype(np.float32),
                         "lpy": rng.uniform(0.02, 0.08, n).astype(np.float32)})
    info = [{"Height": 512, "Width": 512, "Frames": 1, "Pixelsize": 130}]
    render.render(locs.iloc[:100000], info, oversampling=20, blur_method="gaussian")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        k, img = render.render(locs, info, oversampling=20, blur_method="gaussian")
        ts.append(time.perf_counter() - t0)
    out["config4_render_50M_os20"] = {"seconds": min(ts), "all": ts, "locs_per_s": n / min(ts), "n_in_view": int(k),
                                      "image": list(img.shape)}
    del locs, img
    nf, side = 20000, 4096
    dl, dinfo, truth = testing.synthetic_drift_locs(nf, side, side, n_clusters=2000, locs_per_frame=1000.0, seed=3,
                                                    jitter=0.05, lp=0.05)
    postprocess.undrift(dl[dl["frame"] < 500], [{**dinfo[0], "Frames": 500}], 100, display=False,
                        segmentation_callback=lambda i: None, rcc_callback=lambda i: None)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        drift, und = postprocess.undrift(dl, dinfo, 100, display=False, segmentation_callback=lambda i: None,
                                         rcc_callback=lambda i: None)
        ts.append(time.perf_counter() - t0)
    err = [float(np.abs((drift[c].to_numpy() - drift[c].mean()) - (truth[:, k] - truth[:, k].mean())).max())
           for k, c in enumerate(("x", "y"))]
    out["config5_undrift_20M_200seg_4096"] = {"seconds": min(ts), "all": ts, "n_locs": len(dl), "segments": nf // 100,
                                              "pairs": 19900, "max_abs_drift_error_px_vs_injected": err}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
