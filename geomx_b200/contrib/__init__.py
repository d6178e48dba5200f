"""``mx.contrib`` — experimental front ends (parity: python/mxnet/contrib/__init__.py).

``ndarray`` / ``symbol`` are the contrib operator namespaces, ``autograd`` the legacy autograd API, ``io.DataLoaderIter`` wraps a gluon
DataLoader for Module.fit, ``text`` (vocabulary + token embeddings), ``quantization`` (calibrated int8 / fp8 simulation of symbolic
models and gluon nets), ``tensorboard.LogMetricsCallback``, ``svrg_optimization.SVRGModule``.  ``onnx`` / ``tensorrt`` need packages that
are not part of this image and raise a clear error on use."""
from ..ndarray import contrib as ndarray  # noqa: F401
from ..ndarray import contrib as nd  # noqa: F401
from . import autograd, io, quantization, svrg_optimization, tensorboard, text  # noqa: F401
from . import symbol  # noqa: F401
from . import symbol as sym  # noqa: F401


class _Unavailable:
    def __init__(self, name, need):
        self._name, self._need = name, need

    def __getattr__(self, item):
        raise ImportError("mx.contrib.%s needs the '%s' package, which is not installed in this environment" % (self._name, self._need))


onnx = _Unavailable("onnx", "onnx")
tensorrt = _Unavailable("tensorrt", "tensorrt")
