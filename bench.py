#!/usr/bin/env python
"""bench.py -- BASELINE.json headline metric: 7x7 MLE spot-fits/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the MLE hot path (picasso.gaussmle.gaussmle, method
sigmaxy, eps 1e-3, max_it 100) over one batch of synthetic 7x7 spots
(BASELINE.json configs[1]: 10 M spots per GPU).

  value     whole-job fits/s with the ROIs already resident in HBM
            (CUDA-event timed, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call pb_mle_fit
            (pinned host arrays; H2D of the ROIs and D2H of the results inside
            the timed region)
  roofline  algorithmic HBM bytes (252 B/spot) / kernel time vs the measured
            copy bandwidth (the metric asks for HBM GB/s; the fit is FP64-pipe
            bound, see `compute`)
  e2e_python  the same metric through the signature a picasso caller uses:
            picasso_b200.gaussmle.gaussmle(pageable numpy array)
  cpu_baseline  picasso's OWN numba path (gaussmle_async, threads = min(60,
            0.75 * cores)) from the staged reference module oracle/_ref
            (kind "reference"), with the oracle C port (same arithmetic,
            bit-identical results) beside it; the port alone when oracle/_ref or
            numba is missing (kind "port"); bounded sample
  stages    BASELINE configs 3 / 4 / 5 (fused localize, 50 M-loc render, 200 x 4096^2
            undrift) at this world size through picasso_b200.distributed:
            device-resident and end-to-end seconds, per-phase ms, roofline entry,
            parity vs a 1-GPU run (tools/stage_bench.py)

--impl reference times the reference's CPU implementation on the host cores
(oracle/_ref = picasso's own gaussmle.py staged by oracle/make_ref.py, numba;
else the oracle port) and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOX = 7
EPS = 0.001
MAX_IT = 100
METHOD = "sigmaxy"
BYTES_PER_SPOT = BOX * BOX * 4 + 56   # 196 B ROI read + 56 B results written
# The fit runs as three kernels (csrc/mle_tps.cu): start values, Newton iterations (dominant,
# ~75 % of the step), CRLB + log-likelihood.  Algorithmic bytes of the dominant kernel per spot:
# ROI 196 + start theta 24 read, theta 24 + iterations 4 written.
ITER_BYTES_PER_SPOT = BOX * BOX * 4 + 24 + 28
# dram__bytes_read.sum + dram__bytes_write.sum per spot of the dominant kernel come from the committed
# `ncu --set full` capture (profiles/mle_ncu_compute.json, written by tools/ncu_compute_json.py)
# all three kernels: 3 ROI reads + 2 theta reads + theta x2, iterations x2, crlb, logL written
PIPELINE_DRAM_BYTES_PER_SPOT = 3 * BOX * BOX * 4 + 2 * 24 + (24 + 4) * 2 + 28



def compute_from_profile():
    """Instruction-side view of the dominant kernel from the committed `ncu --set full` capture
    (profiles/mle_ncu_compute.json).  The capture records the SHA-256 of the MLE kernel sources it
    was taken at; `stale` says whether the sources on disk still are those."""
    path = os.path.join(ROOT, "profiles", "mle_ncu_compute.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except Exception as exc:        # noqa: BLE001
        return {"source": None, "note": f"no ncu capture committed ({exc})"}
    d["stale"] = d.get("kernel_source_sha256") != mle_kernel_source_hash()
    return d


def mle_kernel_source_hash():
    import hashlib

    h = hashlib.sha256()
    for name in ("mle_tps.cu", "mle_tps_core.cuh", "erf_table.cuh"):
        with open(os.path.join(ROOT, "picasso_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--spots", type=int, default=10_000_000, help="spots per GPU per step")
    p.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-stages", action="store_true", help="skip the configs 3/4/5 stage block")
    p.add_argument("--stages", default="localize,render,undrift")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


NVLINK_WHY = []     # why the counters could not be read (reported in the JSON line)


def nvlink_counters(index):
    """(tx_bytes, rx_bytes) moved over all NVLink ports of GPU `index` since driver load: NVML field values
    NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX (payload KiB, summed over the links with scope id
    UINT_MAX), else the per-link lines of `nvidia-smi nvlink -gt d`.  None when neither is exposed."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        if all(v.nvmlReturn == 0 for v in vals):
            return tuple(int(v.value.ullVal) * 1024 for v in vals)
        NVLINK_WHY.append("NVML field values: nvmlReturn " + "/".join(str(v.nvmlReturn) for v in vals))
    except Exception as e:      # noqa: BLE001
        NVLINK_WHY.append(f"NVML: {type(e).__name__}: {e}")
    try:
        import re

        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True,
                             text=True, timeout=20).stdout
        tx = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
        rx = [int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
        if tx and rx:
            return sum(tx) * 1024, sum(rx) * 1024
        NVLINK_WHY.append("nvidia-smi nvlink -gt d: " + (out.strip().splitlines()[-1][:120] if out.strip() else "no output"))
    except Exception as e:      # noqa: BLE001
        NVLINK_WHY.append(f"nvidia-smi nvlink: {type(e).__name__}: {e}")
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cpu_info():
    """Threads the CPU arm can really use: affinity mask and cgroup quota, not just cpu_count."""
    n = os.cpu_count() or 1
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = n
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        pass
    usable = aff if quota is None else max(1, min(aff, int(math.ceil(quota))))
    return {"cpu_count": n, "affinity": aff, "cgroup_quota_cpus": quota, "threads_used": usable}


_BEST_THREADS = {}


def cpu_oracle_rate(seconds: float, threads: int):
    """Time the CPU oracle on a bounded sample of the same workload.  Oversubscribed or
    quota-limited hosts do not always run fastest with one thread per logical CPU, so a
    short scan (threads, threads/2, threads/4 ...) picks the best count first."""
    import oracle
    from picasso_b200 import testing

    oracle.build()
    if threads not in _BEST_THREADS:
        probe = testing.synthetic_spots(2000 * min(threads, 32), BOX, seed=999)
        best, best_rate, t = threads, 0.0, threads
        while t >= 1:
            oracle.gaussmle(probe[: 64 * t], EPS, MAX_IT, METHOD, nthreads=t)
            t0 = time.perf_counter()
            oracle.gaussmle(probe, EPS, MAX_IT, METHOD, nthreads=t)
            rate = len(probe) / (time.perf_counter() - t0)
            if rate > best_rate:
                best, best_rate = t, rate
            if t == 1:
                break
            t = max(1, t // 2)
        _BEST_THREADS[threads] = best
    use = _BEST_THREADS[threads]
    per_call = 4000 * use
    spots = testing.synthetic_spots(per_call, BOX, seed=12345)
    done, t0 = 0, time.perf_counter()
    while True:
        oracle.gaussmle(spots, EPS, MAX_IT, METHOD, nthreads=use)
        done += per_call
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, done, el, use


_REF = {}


def numba_reference():
    """picasso's own gaussmle module staged under oracle/_ref (None when unavailable)."""
    if "gm" not in _REF:
        try:
            from oracle import make_ref

            make_ref.build()          # no-op on the GPU box (no /root/reference there)
            gm = make_ref.import_gaussmle()
            from picasso_b200 import testing

            t0 = time.perf_counter()
            gm.gaussmle(testing.synthetic_spots(64, BOX, seed=1), EPS, MAX_IT, METHOD)      # JIT warm-up
            _REF["jit_s"] = time.perf_counter() - t0
            _REF["gm"] = gm
        except Exception as exc:      # noqa: BLE001
            print(f"[bench] numba reference unavailable: {exc}", file=sys.stderr)
            _REF["gm"] = None
            _REF["why"] = str(exc)
    return _REF["gm"]


def cpu_numba_rate(seconds: float):
    """Time the reference's gaussmle_async (numba, nogil threads = min(60, 0.75 * cores),
    gaussmle.py:478-530) to completion on a bounded sample of the same workload."""
    import multiprocessing

    from picasso_b200 import testing

    gm = numba_reference()
    workers = min(60, max(1, int(0.75 * multiprocessing.cpu_count())))

    def run(spots):
        t0 = time.perf_counter()
        cur, th, cr, ll, it = gm.gaussmle_async(spots, EPS, MAX_IT, METHOD)
        while not bool((it != 0).all()):           # every fit takes >= 1 iteration
            time.sleep(0.002)
        return time.perf_counter() - t0

    probe = testing.synthetic_spots(4000, BOX, seed=777)
    rate = len(probe) / run(probe)
    per_call = int(max(4000, min(2_000_000, rate * max(seconds, 1.0) / 3)))
    spots = testing.synthetic_spots(per_call, BOX, seed=12345)
    done, el = 0, 0.0
    while el < seconds:
        el += run(spots)
        done += per_call
    return done / el, done, el, workers


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation on the host cores -- picasso's own
    numba gaussmle_async from oracle/_ref when it is staged, else the oracle C port."""
    if rank != 0:
        return
    cpu = host_cpu_info()
    threads = cpu["threads_used"]
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    use_numba = numba_reference() is not None
    rates, n_done, t_tot = [], 0, 0.0
    for i in range(args.warmup + args.steps):
        if use_numba:
            r, d, el, used = cpu_numba_rate(per_step_budget)
        else:
            r, d, el, used = cpu_oracle_rate(per_step_budget, threads)
        if i >= args.warmup:
            rates.append(r)
            n_done += d
            t_tot += el
    value = n_done / t_tot
    if use_numba:
        kind = "reference"
        sample = (f"{n_done} spots in {t_tot:.1f} s: picasso.gaussmle.gaussmle_async (the reference's own "
                  f"numba code from oracle/_ref, {used} nogil threads = min(60, 0.75 * cpu_count), "
                  f"after a {_REF.get('jit_s', 0):.0f} s JIT warm-up)")
    else:
        kind = "port"
        sample = (f"{n_done} spots in {t_tot:.1f} s, oracle C port of picasso.gaussmle._mlefit_sigmaxy "
                  "(bit-identical to the numba reference), pthreads; oracle/_ref unavailable: "
                  + _REF.get("why", "?"))
    line = {
        "impl": "reference", "metric": "MLE spot-fits/sec (7x7 ROI)", "value": value,
        "unit": "fits/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 math / f32 state",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 7x7 MLE sigmaxy eps=1e-3 max_it=100 (bounded CPU sample)",
                   "box": BOX, "method": METHOD},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": used, "kind": kind,
                         "host": cpu, "sample": sample},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def gen_spots_device(torch, n, box, seed, device):
    """Config-2 distribution generated on the GPU (torch) in chunks -> f32 (n, box, box)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, box, box), dtype=torch.float32, device=device)
    c = box // 2
    ii = torch.arange(box, device=device, dtype=torch.float64)[None, :]
    step = 1_000_000
    for lo in range(0, n, step):
        m = min(step, n - lo)
        u = torch.rand((6, m), generator=g, device=device, dtype=torch.float64)
        x0 = c - 0.5 + u[0]
        y0 = c - 0.5 + u[1]
        sx = 0.9 + 0.4 * u[2]
        sy = 0.9 + 0.4 * u[3]
        ph = 500 + 4500 * u[4]
        bg = 5 + 25 * u[5]

        def dE(mu, s):
            a = (ii - mu[:, None] + 0.5) / (math.sqrt(2) * s[:, None])
            b = (ii - mu[:, None] - 0.5) / (math.sqrt(2) * s[:, None])
            return 0.5 * (torch.erf(a) - torch.erf(b))

        ex, ey = dE(x0, sx), dE(y0, sy)
        mu = ph[:, None, None] * ey[:, :, None] * ex[:, None, :] + bg[:, None, None]
        out[lo:lo + m] = torch.poisson(mu.float(), generator=g)
    return out


_REAL_STDOUT = None


def _emit(line: dict):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr for the whole run."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    # keep stdout to the single JSON line: fd 1 -> stderr until the line is emitted
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from picasso_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    _lib.require_gpu()
    if world > 1:
        if "PB_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["PB_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)

    n = args.spots
    spots = gen_spots_device(torch, n, BOX, 1000 + rank, dev)
    # outputs of one rank live in one flat buffer so the multi-GPU gather is a
    # single NCCL all-gather: [thetas 6n | crlbs 6n | logliks n | iterations n].
    # Two such buffers alternate: the all-gather of step i (async, NCCL stream) overlaps the
    # fit of step i+1, which writes the other buffer (the kernel's dynamic tile scheduler
    # absorbs the SMs NCCL borrows); a buffer is reused only after its gather completed.
    flats = [torch.empty(14 * n, dtype=torch.float32, device=dev) for _ in range(2 if world > 1 else 1)]
    works = [None, None]
    stream = torch.cuda.current_stream()
    # The gather itself: "p2p" (default) = copy-engine peer writes into IPC-shared buffers
    # (picasso_b200.distributed.PeerGather: no SM taken from the running fit, no host blocking),
    # "nccl" = one NCCL all-gather per step (PB_GATHER=nccl).  If the peer mapping cannot be set up
    # on every rank the run falls back to NCCL.
    # Default ("auto"): at 2 ranks the all-gather is FUSED into the fit kernel (multimem.st through an NVSwitch
    # multicast mapping); from 4 ranks on it is done by copy-engine peer writes behind the next step's fit --
    # measured at 8 ranks (profiles/r02_summary.md): fused 20.9 ms/step (every GPU would have to take in 3.9 GB
    # within the 3.3 ms of the CRLB kernel), peer copies 18.4 ms/step.
    gather_mode = os.environ.get("PB_GATHER", "auto") if world > 1 else "none"
    if gather_mode == "auto":
        gather_mode = "nvls" if world <= 2 else "p2p"
    gathered, pg, mcb = None, None, None
    nvls_variant = None
    if world > 1 and gather_mode.startswith("nvls"):
        nvls_variant = {"nvls": "fused", "nvls-split": "split", "nvls-split-kernel": "split-kernel",
                        "nvls-ce": "ce"}.get(gather_mode, "fused")
        gather_mode = "nvls"
    if world > 1 and gather_mode == "nvls":
        # fused fit + all-gather: the CRLB kernel stores every spot's results through an NVSwitch multicast
        # mapping into the gather buffers of all ranks (picasso_b200.distributed.MulticastBuffer)
        from picasso_b200.distributed import MulticastBuffer
        import ctypes as C
        lib.pb_mle_fit_gather_dev.argtypes = [C.c_size_t, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int] + \
            [C.c_void_p] * 7
        lib.pb_mle_fit_gather_dev.restype = C.c_int
        lib.pb_mle_fit_gather_mode.argtypes = [C.c_int]
        try:
            mcb = MulticastBuffer(dist, torch, 14 * n * 4, dev)
            _lib.check(lib.pb_mle_fit_gather_mode(2 if nvls_variant.startswith("split") else 1))
            mc_side = torch.cuda.Stream(dev, priority=-1)       # copies of the CRLB half behind the next step
        except Exception as exc:      # noqa: BLE001
            print(f"[bench] rank {rank}: multicast gather unavailable ({exc}); using peer copies", file=sys.stderr)
            gather_mode = "p2p"
    if world > 1 and gather_mode == "p2p":
        from picasso_b200.distributed import PeerGather
        ok = torch.ones(1, device=dev)
        try:
            pg = PeerGather(dist, torch, 14 * n * 4, dev)
        except Exception as exc:      # noqa: BLE001
            print(f"[bench] rank {rank}: peer gather unavailable ({exc}); using NCCL", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            if pg is not None:
                pg.close()
            pg, gather_mode = None, "nccl"
    gathered_parts = {}
    if world > 1 and gather_mode not in ("p2p", "nvls"):
        gather_mode = "nccl"
        gathered = torch.empty(14 * n * world, dtype=torch.float32, device=dev)

    class _Pending:
        """Uniform handle: NCCL work object (host wait) or peer-gather events (device wait)."""
        def __init__(self, work=None, events=None):
            self.work, self.events = work, events

        def wait(self):
            if self.work is not None:
                self.work.wait()
            if self.events is not None:
                for ev in self.events:
                    stream.wait_event(ev)

    def launch_gather(part, offset_bytes, b=0):
        if gather_mode == "p2p":
            if p2p_early:
                m_ = part.numel() // 14
                evs = pg.gather_async(part[: 6 * m_], stream, offset_bytes, after_event=phase_events[b])
                evs += pg.gather_async(part[13 * m_:], stream, offset_bytes + 13 * m_ * 4, after_event=phase_events[b])
                evs += pg.gather_async(part[6 * m_: 13 * m_], stream, offset_bytes + 6 * m_ * 4)
                return _Pending(events=evs)
            return _Pending(events=pg.gather_async(part, stream, offset_bytes))
        # NCCL: one all-gather per part into the part's own gather buffer (part-major as well)
        return _Pending(work=dist.all_gather_into_tensor(
            gathered_parts[offset_bytes], part, async_op=True))

    # PB_BENCH_PARTS > 1 fits the batch in sub-batches and gathers each as soon as it is done
    # (part-major block: per part [thetas 6m | crlbs 6m | logliks m | iterations m]).  Measured:
    # every extra launch of the persistent iteration kernel costs ~1 ms of drain phase (4 parts:
    # 21.9 ms instead of 18.0 ms per 10 M spots), far more than the exposed tail of the last
    # gather it would hide -- so the default stays one launch per step.
    parts = int(os.environ.get("PB_BENCH_PARTS", "1"))
    parts = max(1, min(parts, 64))
    pb = [((n * q) // parts) // 4096 * 4096 for q in range(parts)] + [n]     # part edges, 4096-spot aligned

    # p2p: thetas + iterations are final after the iteration kernel -- the library records a phase event
    # there, and that half of the block leaves while the CRLB kernel still runs (PB_P2P_EARLY=0: one copy
    # of the whole block after the fit)
    p2p_early = os.environ.get("PB_P2P_EARLY", "1") != "0" and parts == 1
    phase_events = []
    if gather_mode == "p2p" and p2p_early:
        import ctypes as C
        lib.pb_mle_set_phase_event.argtypes = [C.c_void_p]
        for _ in range(2):
            ev = torch.cuda.Event()
            ev.record(stream)                    # creates the underlying cudaEvent
            phase_events.append(ev)


    def views(flat, q=None):
        if q is None:                      # all parts: only meaningful for parts == 1
            q = 0
        lo, hi = pb[q], pb[q + 1]
        m = hi - lo
        base = flat[14 * lo: 14 * hi]
        return (base[: 6 * m], base[6 * m: 12 * m], base[12 * m: 13 * m], base[13 * m:].view(torch.int32), base, lo, m)

    def fit_and_gather(flat, b):
        pend = []
        for q in range(parts):
            th, cr, ll, it, base, lo, m = views(flat, q)
            if mcb is not None:
                if nvls_variant == "ce":
                    # unfused reference point: plain fit, then ONE copy-engine copy of the block through the
                    # multicast mapping on a side stream (overlaps the next step)
                    _lib.check(lib.pb_mle_fit_dev(m, BOX, spots[lo:].data_ptr(), EPS, MAX_IT, 1, th.data_ptr(),
                                                  cr.data_ptr(), ll.data_ptr(), it.data_ptr(), None,
                                                  stream.cuda_stream))
                    ready = torch.cuda.Event(); ready.record(stream)
                    mc_side.wait_event(ready)
                    _lib.check(lib.pb_copy_d2d_async(mcb.block_mc_ptr() + 14 * lo * 4, base.data_ptr(), 14 * m * 4,
                                                     mc_side.cuda_stream))
                    done = torch.cuda.Event(); done.record(mc_side)
                    pend.append(_Pending(events=[done]))
                    continue
                _lib.check(lib.pb_mle_fit_gather_dev(m, BOX, spots[lo:].data_ptr(), EPS, MAX_IT, 1, th.data_ptr(),
                                                     cr.data_ptr(), ll.data_ptr(), it.data_ptr(), None,
                                                     mcb.block_mc_ptr() + 14 * lo * 4, stream.cuda_stream))
                if nvls_variant.startswith("split"):
                    # theta + iterations left from the iteration kernel; the CRLB / logL half ([6m, 13m) floats of
                    # the block) follows through the mapping behind the next step: copy engine or a few CTAs
                    ready = torch.cuda.Event(); ready.record(stream)
                    mc_side.wait_event(ready)
                    dst = mcb.block_mc_ptr() + (14 * lo + 6 * m) * 4
                    if nvls_variant == "split":
                        _lib.check(lib.pb_copy_d2d_async(dst, cr.data_ptr(), 7 * m * 4, mc_side.cuda_stream))
                    else:
                        _lib.check(lib.pb_mc_copy_async(dst, cr.data_ptr(), 7 * m * 4, 16, mc_side.cuda_stream))
                    done = torch.cuda.Event(); done.record(mc_side)
                    pend.append(_Pending(events=[done]))
                continue
            if phase_events:
                _lib.check(lib.pb_mle_set_phase_event(phase_events[b].cuda_event))
            _lib.check(lib.pb_mle_fit_dev(m, BOX, spots[lo:].data_ptr(), EPS, MAX_IT, 1, th.data_ptr(),
                                          cr.data_ptr(), ll.data_ptr(), it.data_ptr(), None,
                                          stream.cuda_stream))
            if phase_events:
                _lib.check(lib.pb_mle_set_phase_event(None))
            if world > 1:
                pend.append(launch_gather(base, 14 * lo * 4, b))
        return pend

    if gather_mode == "nccl":
        # NCCL receives each part in its own buffer [world x part]; offsets key the buffers
        off = 0
        for q in range(parts):
            m = pb[q + 1] - pb[q]
            gathered_parts[14 * pb[q] * 4] = gathered[off: off + 14 * m * world]
            off += 14 * m * world

    def wait_all(pend):
        for p_ in pend or ():
            p_.wait()

    def step(i):
        b = i % len(flats)
        wait_all(works[b])
        works[b] = fit_and_gather(flats[b], b)

    def drain():
        for b in range(2):
            wait_all(works[b])
            works[b] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    barrier()
    _lib.check(lib.pb_mle_profile(1))     # CUDA events around the three kernels of each call
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    nvl0 = nvlink_counters(local) if world > 1 else None     # (before the barrier: the ranks start the timed loop together)
    barrier()
    e0.record()
    for i in range(args.steps):
        b = i % len(flats)
        wait_all(works[b])
        k0[i].record()
        works[b] = fit_and_gather(flats[b], b)
        k1[i].record()
    drain()          # every step's gather has left this rank inside the timed region (max over ranks)
    e1.record()
    barrier()
    nvl1 = nvlink_counters(local) if world > 1 else None
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    gather_ok, verified_steps = None, 0
    nvlink = None
    if world > 1:
        # NVLink payload bytes of this rank's GPU over the timed region (counters read outside it), max over ranks
        ok = nvl0 is not None and nvl1 is not None
        t = torch.tensor([float(nvl1[0] - nvl0[0]) if ok else -1.0, float(nvl1[1] - nvl0[1]) if ok else -1.0],
                         dtype=torch.float64, device=dev)
        lo = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        if float(lo.min().item()) >= 0.0:
            nvlink = {"tx_bytes_per_step": float(t[0].item()) / args.steps,
                      "rx_bytes_per_step": float(t[1].item()) / args.steps,
                      "algorithmic_rx_bytes_per_step": (world - 1) * 56 * n,
                      "source": "NVML NVLINK_THROUGHPUT_DATA_TX / _RX (payload, all links of the rank's GPU; what "
                                "`nvidia-smi nvlink -gt d` prints), read before and after the timed region, max over ranks"}
        else:
            nvlink = {"unavailable": "; ".join(dict.fromkeys(NVLINK_WHY)) or "counters not exposed on another rank"}

    def gathered_equals_nccl(b):
        """The gathered array of the step that wrote flats[b], element for element, against an NCCL
        all-gather of the same blocks (all ranks)."""
        ref = torch.empty(14 * n * world, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ref, flats[b])
        if mcb is not None:
            full = mcb.local(torch.int32)
        elif pg is not None:
            pg.finish()
            full = pg.to_tensor(torch.int32)
        else:
            full = torch.cat([gp.view(torch.int32).view(world, -1) for gp in gathered_parts.values()], 1).reshape(-1)
        same = torch.tensor([1.0 if torch.equal(full.view(torch.int32), ref.view(torch.int32)) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        return bool(same.item() > 0.5)

    if world > 1:
        # the gathered array of the LAST timed step is complete and correct on every rank ...
        gather_ok = gathered_equals_nccl((args.steps - 1) % len(flats))
        # ... and so is every step of an untimed verification pass (same launches as the timed loop): the
        # data of each step are perturbed (one spot block is rotated) so that a stale buffer cannot pass
        if parts == 1:
            for v in range(args.steps):
                spots[:4096] = torch.roll(spots[:4096], shifts=v + 1, dims=0)
                b = v % len(flats)
                wait_all(works[b])
                works[b] = fit_and_gather(flats[b], b)
                drain()
                barrier()
                if not gathered_equals_nccl(b):
                    gather_ok = False
                    break
                verified_steps += 1
                barrier()
            works[0] = fit_and_gather(flats[0], 0)      # flats[0] again holds the fit of the final `spots`
            drain()
            barrier()
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in zip(k0, k1)]))
    k3 = np.zeros(3, np.float32)          # {start values, iterations, CRLB} of the last step
    if lib.pb_mle_get_impl() != 0:
        _lib.check(lib.pb_mle_profile_read(k3.ctypes.data))
    _lib.check(lib.pb_mle_profile(0))
    t = torch.tensor([ms_total, ms_kernel, *k3.tolist()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_kernel, ms_init, ms_iter, ms_crlb = t.tolist()
    it = torch.cat([views(flats[0], q)[3] for q in range(parts)])      # iterations in spot order
    mean_it = float(it.float().mean().item())
    value = n * world * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the host-buffer C ABI (pinned host arrays) ----
    e2e = None
    if not args.no_e2e:
        ne = n
        hs = _lib.PinnedArray((ne, BOX, BOX), np.float32)
        hth = _lib.PinnedArray((ne, 6), np.float32)
        hcr = _lib.PinnedArray((ne, 6), np.float32)
        hll = _lib.PinnedArray((ne,), np.float32)
        hit = _lib.PinnedArray((ne,), np.int32)
        torch.from_numpy(hs.array).copy_(spots.cpu())

        def e2e_step():
            _lib.check(lib.pb_mle_fit(ne, BOX, _lib.ptr(hs.array), EPS, MAX_IT, 1,
                                      _lib.ptr(hth.array), _lib.ptr(hcr.array),
                                      _lib.ptr(hll.array), _lib.ptr(hit.array), None, None))

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el = tt.item()
        # result sanity: e2e output equals the device-resident run
        same = bool(np.array_equal(hit.array, it.cpu().numpy()))
        e2e = {"value": ne * world * args.steps / el, "unit": "fits/s",
               "h2d_bytes_per_step": ne * BOX * BOX * 4, "d2h_bytes_per_step": ne * 56,
               "matches_device_run": same}
        for h in (hs, hth, hcr, hll, hit):
            h.free()

        # ---- the signature a picasso caller uses: gaussmle.gaussmle(pageable ndarray) ----
        from picasso_b200 import gaussmle as pb_gaussmle

        hp = spots.cpu().numpy()              # ordinary pageable memory
        t0 = time.perf_counter()
        pb_gaussmle.gaussmle(hp, EPS, MAX_IT, METHOD)
        first_call_s = time.perf_counter() - t0          # includes pinning the result arrays (pooled afterwards)
        pb_gaussmle.gaussmle(hp, EPS, MAX_IT, METHOD)
        barrier()
        pth = pcr = pll = pit = None
        t0 = time.perf_counter()
        for _ in range(args.steps):
            del pth, pcr, pll, pit            # the caller is done with the previous result: its page-locked
            pth, pcr, pll, pit = pb_gaussmle.gaussmle(hp, EPS, MAX_IT, METHOD)       # blocks return to the pool
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_python = {"value": ne * world * args.steps / tt.item(), "unit": "fits/s",
                      "call": "picasso_b200.gaussmle.gaussmle(spots: pageable float32 ndarray, 0.001, 100, "
                              "'sigmaxy') -> 4 ndarrays (page-locked, pooled)",
                      "h2d_bytes_per_step": ne * BOX * BOX * 4, "d2h_bytes_per_step": ne * 56,
                      "first_call_ms": 1e3 * first_call_s,
                      "matches_device_run": bool(np.array_equal(pit, it.cpu().numpy()))}
        del hp, pth, pcr, pll, pit
    else:
        e2e_python = None

    stages = None
    if not args.no_stages:
        del spots, flats
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import stage_bench

        stages = stage_bench.run_stages(torch, dist if world > 1 else None, rank, world, dev, measured_peaks()[0],
                                        which=[w for w in args.stages.split(",") if w])

    if rank == 0:
        peak, peak_src = measured_peaks()
        comp = compute_from_profile()
        tps = ms_iter > 0
        n_launch = pb[parts] - pb[parts - 1]     # spots of the last pb_mle_fit_dev call (profiled one)
        if tps:
            ach = ITER_BYTES_PER_SPOT * n_launch / (ms_iter * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "tps_iter_kernel<7,1,float> (Newton iterations)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": (comp.get("dram_bytes_per_spot") * n_launch
                                if comp.get("dram_bytes_per_spot") else None),
                    "spots_per_launch": n_launch,
                    "traffic_source": "ncu --set full, profiles/mle_ncu_compute.json (bytes/spot x spots per launch)"
                                      + (" -- STALE capture" if comp.get("stale") else ""),
                    "algorithmic_bytes": ITER_BYTES_PER_SPOT * n_launch, "peak_source": peak_src,
                    "kernel_ms": ms_iter,
                    "step_kernels_ms": {"tps_init_kernel": ms_init, "tps_iter_kernel": ms_iter,
                                        "tps_crlb_kernel": ms_crlb, "step": ms_kernel},
                    "step_algorithmic_GBs": BYTES_PER_SPOT * n / (ms_kernel * 1e-3) / 1e9,
                    "step_dram_bytes": PIPELINE_DRAM_BYTES_PER_SPOT * n,
                    "note": "algorithmic 248 B/spot for the iteration kernel (252 B/spot for the "
                            "whole fit; the three kernels together move 720 B/spot = "
                            f"{PIPELINE_DRAM_BYTES_PER_SPOT * n / (ms_kernel * 1e-3) / 1e9 / peak:.1%} of "
                            "HBM peak). The fit is instruction-issue / FP32-pipe bound, not HBM "
                            "bound (SURVEY.md 8d) -- see `compute` and DESIGN.md 5.1"}
        else:
            ach = BYTES_PER_SPOT * n / (ms_kernel * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "mle_fit_kernel<7,8,1> (lane-group)", "achieved": ach,
                    "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": 238.6 * n,
                    "algorithmic_bytes": BYTES_PER_SPOT * n, "peak_source": peak_src,
                    "kernel_ms": ms_kernel}
        line = {
            "metric": "MLE spot-fits/sec (7x7 ROI)", "value": value, "unit": "fits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64 PSF (table erf) / f32 derivative factors and pixel sums with float-float residual / f32 state",
            "data": "synthetic",
            "config": {"workload": "configs[1]: 10M synthetic 7x7 spots/GPU, gaussmle sigmaxy "
                                   "eps=1e-3 max_it=100", "box": BOX, "method": METHOD,
                       "spots_per_gpu": n, "mean_iterations": mean_it,
                       "mle_impl": int(lib.pb_mle_get_impl()),
                       "l2": "input 1.96 GB per step >> 126 MB L2 (no flush needed)",
                       "parallelism": f"spots sharded by index over {world} GPU(s)"
                                      + ({"nccl": "; one NCCL all-gather of the packed outputs per step, "
                                                  "overlapped with the next step's fit",
                                          "nvls": "; all-gather of the packed outputs FUSED into the fit: the kernel that "
                                                  "finishes a spot stores its 56 B of results through an NVSwitch "
                                                  "multicast mapping (multimem.st) into the gather buffers of all "
                                                  "ranks (PB_GATHER=p2p / nccl select the unfused variants)",
                                          "p2p": "; all-gather of the packed outputs per step by copy-engine "
                                                 "peer writes into IPC-shared buffers over NVLink "
                                                 "(PeerGather; PB_GATHER=nccl selects one NCCL all-gather), "
                                                 "overlapped with the next step's fit"}.get(gather_mode, "")),
                       "gather": gather_mode + (f" ({nvls_variant})" if nvls_variant else ""),
                       "gather_verified": gather_ok,
                       "gather_verified_steps": verified_steps, "parts_per_step": parts},
            "roofline": roof,
            # instruction-side view of the dominant kernel from the committed ncu --set full
            # capture (profiles/mle_ncu_compute.json): what actually bounds the fit
            "compute": comp,
            "clocks": clocks, "gpu_launches": launches,
        }
        if nvlink is not None:
            line["nvlink"] = nvlink
        if e2e is not None:
            line["e2e"] = e2e
        if e2e_python is not None:
            line["e2e_python"] = e2e_python
        if stages is not None:
            line["stages"] = stages
        if not args.no_cpu:
            cpu = host_cpu_info()
            threads = cpu["threads_used"]
            r, d, el, used = cpu_oracle_rate(args.cpu_seconds / 2, threads)
            port = {"value": r, "unit": "fits/s", "cores": used, "kind": "port",
                    "sample": f"{d} spots in {el:.1f} s (same distribution), oracle C port of "
                              "picasso.gaussmle._mlefit_sigmaxy, bit-identical to the numba reference"}
            if numba_reference() is not None:
                r2, d2, el2, used2 = cpu_numba_rate(args.cpu_seconds / 2)
                line["cpu_baseline"] = {
                    "value": r2, "unit": "fits/s", "cores": used2, "kind": "reference", "host": cpu,
                    "sample": f"{d2} spots in {el2:.1f} s (same distribution): picasso.gaussmle.gaussmle_async, "
                              f"the reference's own numba code staged in oracle/_ref, {used2} nogil threads = "
                              "min(60, 0.75 * cpu_count) (gaussmle.py:503), after a JIT warm-up call",
                    "port": port}
            else:
                port["host"] = cpu
                port["numba_reference_unavailable"] = _REF.get("why", "?")
                line["cpu_baseline"] = port
        _emit(line)
    if pg is not None:
        pg.close()
    if mcb is not None:
        mcb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
