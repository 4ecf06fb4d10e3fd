"""profiles/mle_ncu_compute.json from an `ncu --set full` capture of the MLE kernels.

    python tools/ncu_compute_json.py gpurun_out/<capture>.ncu-rep <spots per launch> [out.json]

Reads the raw page of the report (`ncu -i ... --page raw --csv`), keeps the tps_* kernels and writes
per kernel: duration, registers, warp instructions (total and per spot), issue / pipe utilisation,
DRAM bytes (total and per spot) -- plus the SHA-256 of the kernel sources the capture was taken at
(csrc/mle_tps.cu + mle_tps_core.cuh).  bench.py prints this block as `compute` and marks it
`stale: true` when the sources on disk no longer hash to the recorded value.  Numbers under ncu are
never bench values; they explain them."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = {
    "duration_ms": "gpu__time_duration.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "warp_instructions": "smsp__inst_executed.sum",
    "threads_per_warp_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fp64_pipe_busy_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "fma_pipe_busy_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "xu_pipe_busy_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "alu_pipe_busy_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu_pipe_busy_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}


def main():
    rep, spots = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "mle_ncu_compute.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names, units = rows[0], rows[1]
    kcol = names.index("Kernel Name")
    kernels = {}
    for r in rows[2:]:
        kn = r[kcol]
        short = next((k for k in ("tps_iter_kernel", "tps_crlb_kernel", "tps_init_kernel") if k in kn), None)
        if short is None or short in kernels:
            continue
        d = {}
        for key, metric in METRICS.items():
            if metric not in names:
                continue
            i = names.index(metric)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if key.endswith("_bytes") or key == "duration_ms":
                v *= UNIT.get(u, 1.0)
            d[key] = v
        if "warp_instructions" in d:
            d["warp_instructions_per_spot"] = d["warp_instructions"] / spots
        if "dram_read_bytes" in d:
            d["dram_bytes_per_spot"] = (d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)) / spots
        kernels[short] = d
    import bench

    from picasso_b200 import build as pb_build

    doc = {"source": f"ncu --set full --clock-control none, {os.path.basename(rep)}, {spots} spots per launch "
                     "(7x7, sigmaxy, eps 1e-3, max_it 100)",
           "kernel_source_sha256": bench.mle_kernel_source_hash(),
           "library_source_sha256": pb_build.source_hash(),
           "spots_per_launch": spots, "kernels": kernels}
    dom = kernels.get("tps_iter_kernel", {})
    for k in ("issue_slots_busy_pct", "fp64_pipe_busy_pct", "fma_pipe_busy_pct", "xu_pipe_busy_pct",
              "warp_instructions_per_spot", "dram_bytes_per_spot"):
        if k in dom:
            doc[k] = dom[k]
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc)[:600])


if __name__ == "__main__":
    main()
