"""Import the *real* picasso reference from /root/reference with its absent
GUI / file-format dependencies mocked (SURVEY.md section 8c recipe).

This module is only used in the build container (where /root/reference is
mounted) by ``tools/gen_golden.py`` to generate the committed golden vectors
under ``tests/golden/``.  Nothing in ``tests/``, ``bench.py`` or the product
imports it at run time: /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("PICASSO_REFERENCE", "/root/reference")

_MOCKS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
    "matplotlib.backends", "matplotlib.backends.backend_qt5agg",
    "matplotlib.backends.backend_qtagg", "matplotlib.figure",
    "matplotlib.patches", "matplotlib.widgets", "matplotlib.gridspec",
    "mpl_toolkits", "mpl_toolkits.mplot3d",
    "PyQt6", "PyQt6.QtCore", "PyQt6.QtGui", "PyQt6.QtSvg",
    "playsound3", "h5py", "tables", "nd2", "tifffile", "dask", "dask.array",
    "sqlalchemy", "imageio", "imageio.v2", "statsmodels",
    "statsmodels.nonparametric", "statsmodels.nonparametric.smoothers_lowess",
    "streamlit", "hdf5plugin", "PyImarisWriter", "lmfit",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "picasso"))


def install_mocks() -> None:
    for name in _MOCKS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock(name=name)
    # lib.py subclasses these Qt widgets at import time -> need real classes
    if not isinstance(sys.modules.get("PyQt6.QtWidgets"), types.ModuleType) or \
            isinstance(sys.modules.get("PyQt6.QtWidgets"), MagicMock):
        qtw = types.ModuleType("PyQt6.QtWidgets")

        class _Dummy:  # minimal stand-in base class
            def __init__(self, *a, **k):
                pass

        def __getattr__(name):  # any widget name resolves to a dummy class
            cls = type(name, (_Dummy,), {})
            setattr(qtw, name, cls)
            return cls

        qtw.__getattr__ = __getattr__
        sys.modules["PyQt6.QtWidgets"] = qtw
        sys.modules["PyQt6"].QtWidgets = qtw


def import_reference():
    """Return the dict of reference modules (picasso.gaussmle, ...)."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    install_mocks()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    mods = {}
    for m in ("lib", "io", "gaussmle", "gausslq", "localize", "render",
              "imageprocess", "postprocess", "zfit", "aim"):
        mods[m] = importlib.import_module(f"picasso.{m}")
    return mods


if __name__ == "__main__":
    ms = import_reference()
    print({k: v.__file__ for k, v in ms.items()})
