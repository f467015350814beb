"""Loader for the in-tree native library ``cpd_b200/_C*.so``.

The library is built ahead of time by ``python -m cpd_b200.build`` (or
``__graft_entry__.build()``).  If it is missing and a CUDA toolchain is present it is
built on first import; if it cannot be loaded the import fails loudly — there is no
silent pure-PyTorch fallback for the kernels.
"""
from __future__ import annotations

import importlib
import os

_C = None
_err = None


def load():
    global _C, _err
    if _C is not None:
        return _C
    try:
        _C = importlib.import_module("cpd_b200._C")
        return _C
    except ImportError as e:  # not built yet
        _err = e
    if os.environ.get("CPD_B200_NO_AUTOBUILD") != "1":
        from . import build as _build

        _build.build(verbose=bool(os.environ.get("CPD_B200_VERBOSE_BUILD")))
        importlib.invalidate_caches()
        _C = importlib.import_module("cpd_b200._C")
        return _C
    raise ImportError(
        "cpd_b200: native library not built (run `python -m cpd_b200.build`): %s" % (_err,))
