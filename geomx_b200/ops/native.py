"""ctypes bridge to the hand-written sm_100a kernels (``geomx_b200/lib/libgeomx_kernels.so``).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m geomx_b200.build`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` and exposes a flat C ABI (raw device pointers +
``cudaStream_t``).  Launches go onto PyTorch's *current* CUDA stream, so they compose with stream capture
(``torch.cuda.graph``) and with torch ops on the same stream.

``require()`` raises when the library is missing on a machine that has a GPU: a silent eager fallback would hide a
broken build (the round-end harness records which in-tree ``.so`` files were actually loaded).
"""
from __future__ import annotations

import ctypes
import os

import torch

_LIB = None
_TRIED = False
_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libgeomx_kernels.so")
launch_count = 0  # number of native kernel launches issued through this bridge (bench.py reports it)


def lib_path():
    return _LIB_PATH


def _load():
    global _LIB, _TRIED
    if _TRIED:
        return _LIB
    _TRIED = True
    if os.path.exists(_LIB_PATH):
        try:
            _LIB = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
            from . import _native_sigs
            _native_sigs.declare(_LIB)
        except OSError as e:  # pragma: no cover
            _LIB = None
            if torch.cuda.is_available():
                raise
    return _LIB


def available() -> bool:
    return torch.cuda.is_available() and _load() is not None


def require():
    if _load() is None:
        raise RuntimeError("native kernel library %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'`" % _LIB_PATH)
    return _LIB


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: CUDA error %d" % (what, rc))


def __getattr__(name):
    # kernel wrappers live in _native_api (kept separate so this loader stays importable without CUDA)
    from . import _native_api
    try:
        return getattr(_native_api, name)
    except AttributeError:
        raise AttributeError("module 'geomx_b200.ops.native' has no attribute %r" % name)
