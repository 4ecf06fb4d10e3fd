"""Two-rank NCCL check of the sharded stages (picasso_b200.distributed) -- runs only where at
least two GPUs are visible (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`);
the gloo world-size-2 tests in test_distributed_cpu.py cover the gathers on CPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_stages_agree_with_one_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "bench_multi.py"),
           "--small"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2
    assert out["localize"]["tables_bit_identical"] is True
    assert out["render"]["n_equal"] is True and out["render"]["max_rel_dev_bright_pixels"] < 1e-4
    assert out["undrift"]["max_abs_drift_dev"] < 1e-5


def test_peer_gather_equals_nccl_all_gather():
    """PeerGather (copy-engine peer writes into IPC-shared buffers) delivers what an NCCL all-gather
    delivers, whole blocks and parts at an offset."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tools", "check_peer_gather.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out == {"world": 2, "peer_gather_equals_nccl": True}


def test_scalable_sharding_agrees_with_one_gpu():
    """Row-band render (all-to-all) and segment-sharded undrift (spectra all-gather, tile-sharded
    pairs) on two ranks against the single-GPU product calls."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "check_sharded_scalable.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["world"] == 2
    for bm in ("gaussian", "gaussian_iso", "None"):
        assert out["render_bands"][bm]["ok"] is True, out["render_bands"][bm]
        assert out["render_bands"][bm]["n"] == out["render_bands"][bm]["n_ref"]
    u = out["undrift_segments"]
    assert u["max_abs_drift_dev_px"] < 1e-5 and u["max_abs_row_dev_px"] < 1e-4
    assert u["rows_total"] == u["rows_ref"]


def test_multicast_fused_gather_equals_nccl():
    """NVSwitch multicast buffer + the fit kernel's fused all-gather (multimem.st) against NCCL."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tools", "check_multicast.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    if not out.get("supported"):
        pytest.skip("no NVSwitch multicast support on this box: " + out.get("why", ""))
    assert out["copy_equals_nccl"] is True
    assert out["fused_fit_gather_equals_nccl"] == [True, True]
