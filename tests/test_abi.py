"""The C-ABI library builds, loads and exports every symbol include/*.h declares.
(No compute calls here -- this runs on the CPU-only build box.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_builds_and_exports_all_declared_symbols():
    from picasso_b200 import build

    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_loader_and_error_convention():
    from picasso_b200 import _lib

    lib = _lib.load()
    assert b"sm_100a" in lib.pb_version()
    assert isinstance(_lib.device_count(), int)
    assert isinstance(_lib.launch_count(), int)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under picasso_b200/ may reference it."""
    pkg = os.path.join(ROOT, "picasso_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), fn
                assert "liboracle" not in text, fn


def test_no_gpu_means_loud_failure():
    """Without a B200 the product raises instead of silently falling back."""
    import numpy as np

    from picasso_b200 import _lib, gaussmle

    if _lib.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.PicassoB200Error):
        gaussmle.gaussmle(np.ones((4, 7, 7), np.float32), 1e-3, 10)
