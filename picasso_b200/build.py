"""Build libpicasso_b200.so in-tree with nvcc for sm_100a.

    python -m picasso_b200.build [--force] [--verbose]

The shared library is the product: a C-ABI (include/picasso_b200.h) with no
torch / Python types in its signatures.  It is git-ignored but travels to the
GPU box with the gpurun snapshot.

Freshness is decided by CONTENT, not by mtime: the SHA-256 of every file under csrc/ and
include/ plus the compiler flags is embedded in the library (``pb_version()`` reports it) and
``needs_build()`` compares the hash of the sources on disk with the one inside the binary -- a
library built from other sources (e.g. one that travelled with a snapshot) is rebuilt.  Objects
are rebuilt individually when the hash of (their source, every header, their flags) changes.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpicasso_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
LIBS = ["-lcufft"]
# per-file flags: the z fit and the linking distances follow the reference's unfused
# floating-point trajectory (numba emits no fused multiply-adds)
FILE_FLAGS = {"zfit.cu": ["-fmad=false"], "link.cu": ["-fmad=false"]}
HASH_MARK = b"PB_SRC_HASH="


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE)) if f.endswith(".h")]
    return hs


def _digest(paths, extra=()):
    h = hashlib.sha256()
    for p in paths:
        h.update(os.path.basename(p).encode() + b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    for e in extra:
        h.update(str(e).encode() + b"\0")
    return h.hexdigest()


def source_hash() -> str:
    """SHA-256 over csrc/*.cu, csrc/*.cuh, include/*.h and the compile flags."""
    return _digest([os.path.join(CSRC, s) for s in _sources()] + _headers(),
                   [*ARCH, *CFLAGS, *LIBS, sorted(FILE_FLAGS.items())])


def library_hash(path: str = LIB):
    """The source hash embedded in a built library (None when absent / not found)."""
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        data = f.read()
    m = re.search(HASH_MARK + rb"([0-9a-f]{64})", data)
    return m.group(1).decode() if m else None


def needs_build() -> bool:
    return library_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    want = source_hash()
    if not force and library_hash() == want:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    headers = _headers()

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        flags = [*ARCH, *CFLAGS, *FILE_FLAGS.get(src, [])]
        if src == "api.cu":
            flags.append(f'-DPB_SOURCE_HASH="{want}"')      # pb_version() reports it
        key = _digest([os.path.join(CSRC, src)] + headers, flags)
        stamp = obj + ".hash"
        if (not force and not verbose and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read().strip() == key):
            return obj
        cmd = [NVCC, *flags, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(key)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-L/usr/local/cuda/lib64",
           "-Xlinker", "-rpath,/usr/local/cuda/lib64", *LIBS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    got = library_hash()
    if got != want:
        raise RuntimeError(f"built library reports source hash {got}, expected {want}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
