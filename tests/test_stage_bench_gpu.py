"""bench.py's `stages` block (tools/stage_bench.py) at reduced sizes on one GPU: every stage runs, reports
its device-resident and end-to-end seconds, a roofline entry and a green parity block -- so that a broken
stage is caught by the test suite and not by an empty block in the driver's bench line."""
import os
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def ctx():
    import torch

    import stage_bench

    from picasso_b200 import _lib

    _lib.check(_lib.load().pb_set_device(0))
    return stage_bench, stage_bench.Ctx(torch, None, 0, 1, torch.device("cuda", 0), 6551.4, repeats=1)


def _common(out):
    assert out["seconds"] > 0 and out["e2e_seconds"] > 0
    r = out["roofline"]
    assert r["bound"] == "hbm" and r["achieved"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9


def test_stage_localize(ctx):
    sb, c = ctx
    out = sb.stage_localize(c, frames=200, Y=256, X=256, chunk=100)
    _common(out)
    assert out["n_localizations"] > 1000
    assert out["parity"]["vs_oracle_identifications_first_8_frames"] is True
    assert out["parity"]["e2e_equals_device_run"] is True


def test_stage_render(ctx):
    sb, c = ctx
    out = sb.stage_render(c, n_total=2_000_000, chunk=125_000)
    _common(out)
    assert out["parity"]["vs_oracle_ok"] is True and out["parity"]["e2e_n_equal"] is True
    assert out["parity"]["e2e_band_max_abs_diff_vs_device_run"] < 1e-4


def test_stage_undrift(ctx):
    sb, c = ctx
    out = sb.stage_undrift(c, n_frames=2000, side=1024, segmentation=100, n_clusters=400)
    _common(out)
    # 20 segments only: the cubic spline through 20 points limits the recovery of the injected drift
    assert max(out["parity"]["max_abs_drift_error_px_vs_injected"]) < 0.1
    assert out["parity"]["e2e_drift_max_abs_diff_vs_device_run"] < 1e-6
    assert {"render_ms", "r2c_ms", "pairs_ms", "peakfit_ms"} <= set(out["phases_ms"])
