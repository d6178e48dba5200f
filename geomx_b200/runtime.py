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


class Feature:
    def __init__(self, name, enabled):
        self.name, self.enabled = name, bool(enabled)

    def __repr__(self):
        return "%s %s" % ("✔" if self.enabled else "✖", self.name)


class Features(dict):
    """Compile-/run-time feature table (``mx.runtime.Features()``; ``is_enabled('CUDA')``).  Names follow libinfo's feature enum where a
    counterpart exists; Blackwell-specific entries are added (TCGEN05, TMA, NVLS_FABRIC)."""

    def __init__(self):
        import torch
        cuda = torch.cuda.is_available()
        kern = bool(glob.glob(os.path.join(lib_dir(), "libgeomx_kernels*.so")))
        feats = {
            "CUDA": cuda, "CUDNN": cuda and torch.backends.cudnn.is_available(), "NCCL": torch.distributed.is_available() and torch.distributed.is_nccl_available(),
            "CUDA_RTC": True, "TENSORRT": False, "CPU_SSE": True, "OPENMP": True, "F16C": True, "BLAS_OPEN": False, "BLAS_MKL": False, "MKLDNN": False,
            "OPENCV": False, "DIST_KVSTORE": available(), "SIGNAL_HANDLER": available(), "PROFILER": available(), "INT64_TENSOR_SIZE": True,
            "NATIVE_KERNELS_SM100A": kern, "TCGEN05": kern, "TMA": kern, "NVLS_FABRIC": kern,
        }
        super().__init__({k: Feature(k, v) for k, v in feats.items()})

    def is_enabled(self, name):
        name = name.upper()
        if name not in self:
            raise RuntimeError("Feature '%s' is unknown, known features are: %s" % (name, list(self)))
        return self[name].enabled


def feature_list():
    return list(Features().values())
