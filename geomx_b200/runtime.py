"""Loader of the native C++ runtime extension ``geomx_b200/lib/_C*.so`` (HiPS transport/servers, codecs, .params IO, profiler, engine,
data IO).  Built in-tree by ``python -m geomx_b200.build``; ``C()`` raises with a build hint when it is missing."""
from __future__ import annotations

import glob
import importlib.util
import os

_mod = None


def lib_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def available():
    return bool(glob.glob(os.path.join(lib_dir(), "_C*.so")))


def C():
    global _mod
    if _mod is None:
        cands = glob.glob(os.path.join(lib_dir(), "_C*.so"))
        if not cands:
            raise RuntimeError("native runtime not built — run `python -m geomx_b200.build`")
        spec = importlib.util.spec_from_file_location("_C", cands[0])
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
