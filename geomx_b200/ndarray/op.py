"""``mx.nd.op`` — every registered operator as a function (reference: ``python/mxnet/ndarray/op.py``, generated from the nnvm registry by
``ndarray/register.py``).  Here operators are ordinary Python functions over torch tensors / hand-written kernels; this module collects them."""
from . import op_lib as _op_lib

__all__ = list(_op_lib.__all__)
for _n in __all__:
    globals()[_n] = getattr(_op_lib, _n)
del _n


def __getattr__(name):            # operators that live in ndarray.py itself (creation, elementwise arithmetic, reductions ...)
    from . import ndarray as _nd
    fn = getattr(_nd, name, None)
    if fn is None or not callable(fn):
        raise AttributeError("mx.nd.op has no operator %r" % name)
    return fn
