"""``mx.libinfo`` — where the native libraries live (parity: python/mxnet/libinfo.py ``find_lib_path`` / ``find_include_path`` /
``__version__``; ``MXNET_LIBRARY_PATH`` overrides the search)."""
from __future__ import annotations

import glob
import os

__version__ = "0.1.0"          # API level: MXNet 1.4 (the GeoMX fork point)


def find_lib_path():
    """Paths of the in-tree native libraries: the pybind runtime ``_C*.so`` (HiPS transport, servers, engine, IO) and the CUDA kernel
    library ``libgeomx_kernels.so``.  Raises when nothing has been built (``python -m geomx_b200.build``)."""
    env = os.environ.get("MXNET_LIBRARY_PATH")
    if env and os.path.isfile(env):
        return [env]
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
    libs = sorted(glob.glob(os.path.join(d, "_C*.so")) + glob.glob(os.path.join(d, "libgeomx_*.so")))
    if not libs:
        raise RuntimeError("Cannot find the native libraries under %s — run `python -m geomx_b200.build`" % d)
    return libs


def find_include_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
