"""oracle/_ref: the REAL reference's MLE module, staged for timing on the GPU box's host cores.

TEST / BENCH INFRASTRUCTURE ONLY.  The Python reference cannot be pip-installed on the GPU box
(no network, GUI dependencies), but its MLE path is one self-contained numba module.  This recipe
copies ``picasso/gaussmle.py`` UNMODIFIED from ``/root/reference`` (build container only) into the
git-ignored ``oracle/_ref/picasso/`` next to a stub ``picasso.lib`` (the module uses ``lib`` only
inside postponed annotations), so that ``bench.py``'s CPU legs can time picasso's own
``gaussmle_async`` (numba, threads = min(60, 0.75 * cores), gaussmle.py:478-530) -- what
``north_star`` asks for -- instead of the C port.  ``oracle/_ref`` is listed in .gitignore (no
reference source enters the history) but not in .gpurunignore (it travels to the GPU box).

    python oracle/make_ref.py          # no-op (exit 0) when /root/reference is absent
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PICASSO_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")

_STUB_INIT = '"""Stub package: only picasso.gaussmle (copied unmodified) is staged here."""\n'
_STUB_LIB = ('"""Stub of picasso.lib: gaussmle.py refers to lib.* only inside postponed annotations\n'
             '(from __future__ import annotations), so no name is ever looked up."""\n')


def available() -> bool:
    """True when the staged reference module exists (on either box)."""
    return os.path.exists(os.path.join(DEST, "picasso", "gaussmle.py"))


def build(force: bool = False) -> bool:
    """Stage oracle/_ref from /root/reference.  Returns True when oracle/_ref is usable."""
    src = os.path.join(REF_ROOT, "picasso", "gaussmle.py")
    if not os.path.exists(src):
        return available()
    pkg = os.path.join(DEST, "picasso")
    dst = os.path.join(pkg, "gaussmle.py")
    if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
        os.makedirs(pkg, exist_ok=True)
        shutil.copyfile(src, dst)
        with open(os.path.join(pkg, "__init__.py"), "w") as f:
            f.write(_STUB_INIT)
        with open(os.path.join(pkg, "lib.py"), "w") as f:
            f.write(_STUB_LIB)
    return True


def import_gaussmle():
    """Import the staged reference module as ``picasso.gaussmle`` (raises ImportError when it is
    not staged or numba is missing)."""
    if not available():
        raise ImportError("oracle/_ref/picasso/gaussmle.py not staged (python oracle/make_ref.py)")
    import numba  # noqa: F401  (fail early and clearly)

    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import importlib

    mod = importlib.import_module("picasso.gaussmle")
    if not os.path.abspath(mod.__file__).startswith(DEST):
        raise ImportError(f"another picasso package shadows oracle/_ref: {mod.__file__}")
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "reference not present, nothing staged")
