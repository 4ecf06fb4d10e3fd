"""Build libpicasso_b200.so in-tree with nvcc for sm_100a.

    python -m picasso_b200.build [--force] [--verbose]

The shared library is the product: a C-ABI (include/picasso_b200.h) with no
torch / Python types in its signatures.  It is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpicasso_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
LIBS = ["-lcufft"]
# per-file flags: the z fit follows the reference's float64 trajectory without fused multiply-adds
FILE_FLAGS = {"zfit.cu": ["-fmad=false"]}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_input():
    t = 0.0
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    t = max(t, os.path.getmtime(os.path.join(HERE, "..", "include", "picasso_b200.h")))
    return t


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_input()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC, *ARCH, *CFLAGS, *FILE_FLAGS.get(src, []), *extra, "-c", os.path.join(CSRC, src),
               "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-L/usr/local/cuda/lib64",
           "-Xlinker", "-rpath,/usr/local/cuda/lib64", *LIBS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
