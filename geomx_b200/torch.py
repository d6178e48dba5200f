"""``mx.th`` / ``mx.torch`` — calling torch functions on NDArrays.

In the reference this module (``python/mxnet/torch.py``) exposes the functions of the optional Lua-Torch plugin (``plugin/torch``).  Here every
NDArray IS a ``torch.Tensor`` underneath, so the bridge is direct: ``mx.th.<fn>(*args)`` calls ``torch.<fn>`` with NDArray arguments unwrapped
(zero copy) and tensor results wrapped back; ``mx.th.to_torch(nd)`` / ``mx.th.from_torch(t)`` convert explicitly."""
import torch as _torch

from .ndarray import NDArray

__all__ = ["to_torch", "from_torch"]


def to_torch(arr):
    """The tensor behind ``arr`` (shares memory)."""
    return arr._t if isinstance(arr, NDArray) else arr


def from_torch(t):
    """Wrap a tensor without copying."""
    return NDArray(t)


def _unwrap(x):
    if isinstance(x, NDArray):
        return x._t
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(v) for v in x)
    return x


def _wrap(x):
    if isinstance(x, _torch.Tensor):
        return NDArray(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_wrap(v) for v in x)
    return x


def __getattr__(name):
    fn = getattr(_torch, name, None)
    if fn is None or not callable(fn):
        raise AttributeError("torch has no function %r" % name)

    def call(*args, **kwargs):
        return _wrap(fn(*_unwrap(args), **{k: _unwrap(v) for k, v in kwargs.items()}))
    call.__name__ = name
    call.__doc__ = "torch.%s on NDArrays (arguments unwrapped, tensor results wrapped)." % name
    return call
