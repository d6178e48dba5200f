"""NDArray: an MXNet-shaped tensor handle over ``torch.Tensor``.

Parity: ``python/mxnet/ndarray/ndarray.py`` (class NDArray :156, ``astype`` :1983,
``asscalar`` :1788, ``copyto`` :2000, ``as_in_context`` :2035, ``attach_grad`` :2099)
and the engine-variable semantics of ``include/mxnet/ndarray.h``.

Design (not a port): the reference gives every NDArray an engine ``Var`` and makes
every op asynchronous through a C++ dependency engine.  On B200 the CUDA stream
already *is* the dependency engine for device work, so an NDArray is a plain
handle ``(_t: torch.Tensor)``.  The one place where the reference's read-after-
write tracking matters is the KVStore: ``pull`` writes into an array
asynchronously and later reads must observe it.  We keep that contract with a
per-array ``_pending`` hook: the KVStore's native scheduler registers a flush
callback on arrays that are targets of queued pulls; touching ``._t`` runs it.
"""
from __future__ import annotations

import numpy as np
import torch

from ..base import MXNetError
from ..context import Context, current_context

__all__ = ["NDArray", "array", "zeros", "ones", "empty", "full", "arange", "zeros_like", "ones_like",
           "waitall", "concat", "stack", "dot", "np_dtype", "torch_dtype", "dtype_name", "from_torch",
           "maximum", "minimum", "sqrt", "square", "abs", "exp", "log", "clip", "sign", "where",
           "softmax", "log_softmax", "relu", "sigmoid", "tanh", "pick", "one_hot", "norm",
           "add_n", "elemwise_add", "moveaxis", "transpose", "reshape", "flatten", "sum", "mean",
           "argmax", "argmin", "max", "min", "topk", "sort", "argsort", "broadcast_to", "tile", "cast"]

_DT = {
    "float32": torch.float32, "float64": torch.float64, "float16": torch.float16,
    "bfloat16": torch.bfloat16, "uint8": torch.uint8, "int8": torch.int8, "int32": torch.int32,
    "int64": torch.int64, "bool": torch.bool,
}
_DT_REV = {v: k for k, v in _DT.items()}


def torch_dtype(dt):
    """Anything dtype-like (str / numpy dtype / torch dtype / python type) -> torch.dtype."""
    if dt is None:
        return torch.float32
    if isinstance(dt, torch.dtype):
        return dt
    if isinstance(dt, str):
        return _DT[dt]
    if dt is float:
        return torch.float32
    if dt is int:
        return torch.int32
    return _DT[np.dtype(dt).name]


def dtype_name(dt) -> str:
    return _DT_REV[torch_dtype(dt)]


def np_dtype(dt):
    n = dtype_name(dt)
    if n == "bfloat16":
        raise MXNetError("bfloat16 has no numpy dtype")
    return np.dtype(n)


def _ctx_of(ctx):
    if ctx is None:
        return current_context()
    if isinstance(ctx, Context):
        return ctx
    if isinstance(ctx, torch.device):
        return Context.from_torch(ctx)
    raise MXNetError("bad context %r" % (ctx,))


def _recording():
    from .. import autograd
    return autograd.is_recording()


class NDArray:
    """Tensor handle.  ``_t`` is the backing torch tensor (property: runs pending KV flush)."""

    __slots__ = ("_data", "_grad", "_grad_req", "_pending", "_ctx_hint", "_stable_grad", "__weakref__")
    __array_priority__ = 1000.0

    def __init__(self, t: torch.Tensor, ctx: Context | None = None):
        self._data = t
        self._grad = None
        self._grad_req = "null"
        self._pending = None
        self._ctx_hint = ctx
        self._stable_grad = False

    # ---- storage access -------------------------------------------------------------------
    @property
    def _t(self) -> torch.Tensor:
        p = self._pending
        if p is not None:
            self._pending = None
            p()
        return self._data

    @_t.setter
    def _t(self, v):
        self._data = v

    # ---- basic properties -----------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._data.shape)

    @property
    def size(self):
        return int(self._data.numel())

    @property
    def ndim(self):
        return self._data.dim()

    @property
    def dtype(self):
        n = _DT_REV[self._data.dtype]
        return np.dtype(n).type if n != "bfloat16" else torch.bfloat16

    @property
    def context(self):
        if self._ctx_hint is not None and self._ctx_hint.torch_device.type == self._data.device.type:
            return self._ctx_hint
        return Context.from_torch(self._data.device)

    ctx = context

    @property
    def stype(self):
        return "default"

    @property
    def T(self):
        return NDArray(self._t.t())

    @property
    def grad(self):
        return self._grad

    # ---- sync -----------------------------------------------------------------------------
    def wait_to_read(self):
        t = self._t
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        return self

    def asnumpy(self):
        t = self._t.detach()
        if t.dtype == torch.bfloat16:
            t = t.float()
        return t.cpu().numpy()

    def asscalar(self):
        if self.size != 1:
            raise ValueError("The current array is not a scalar")
        return self.asnumpy().reshape(-1)[0]

    def item(self):
        return self.asscalar().item()

    def __float__(self):
        return float(self.asscalar())

    def __int__(self):
        return int(self.asscalar())

    def __bool__(self):
        if self.size == 1:
            return bool(self.asscalar())
        raise ValueError("The truth value of an NDArray with multiple elements is ambiguous.")

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return "\n%s\n<NDArray %s @%s>" % (str(self.asnumpy()), "x".join(map(str, self.shape)), self.context)

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i]

    # ---- conversion / movement -------------------------------------------------------------
    def astype(self, dtype, copy=True):
        td = torch_dtype(dtype)
        if not copy and td == self._data.dtype:
            return self
        return NDArray(self._t.to(td, copy=True), self._ctx_hint)

    def as_in_context(self, ctx):
        ctx = _ctx_of(ctx)
        ctx.check_available()
        if ctx.torch_device == self._data.device:
            if ctx.device_typeid in (3, 5) and ctx != self.context:       # same device, other storage kind (pinned / shared host memory)
                return NDArray(ctx.place(self._t.clone()), ctx)
            return self
        return NDArray(ctx.place(self._t.to(ctx.torch_device, non_blocking=True)), ctx)

    def copyto(self, other):
        if isinstance(other, NDArray):
            if other is self:
                return other
            other._t.copy_(self._t, non_blocking=True)
            return other
        ctx = _ctx_of(other)
        ctx.check_available()
        return NDArray(ctx.place(self._t.to(ctx.torch_device, copy=True)), ctx)

    def copy(self):
        return NDArray(self._t.clone(), self._ctx_hint)

    def detach(self):
        return NDArray(self._t.detach(), self._ctx_hint)

    def reshape(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        src = self.shape
        # MXNet's 0 = "copy this dim" convention
        shape = [src[i] if (s == 0 and i < len(src)) else s for i, s in enumerate(shape)]
        return NDArray(self._t.reshape(shape))

    def flatten(self):
        return NDArray(self._t.reshape(self.shape[0], -1))

    def expand_dims(self, axis):
        return NDArray(self._t.unsqueeze(axis))

    def squeeze(self, axis=None):
        return NDArray(self._t.squeeze() if axis is None else self._t.squeeze(axis))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self.ndim)))
        return NDArray(self._t.permute(*axes))

    def broadcast_to(self, shape):
        return NDArray(self._t.expand(*shape))

    def tostype(self, stype):
        if stype in ("row_sparse", "csr"):
            from . import sparse
            return sparse.cast_storage(self, stype)
        if stype != "default":
            raise MXNetError("storage type %s is not supported (default / row_sparse / csr)" % stype)
        return self

    # ---- autograd -------------------------------------------------------------------------
    def attach_grad(self, grad_req="write", stype=None):
        t = self._t
        if not t.is_floating_point():
            raise MXNetError("attach_grad needs a floating array")
        if not t.is_leaf or t.requires_grad is False:
            t = t.detach()
        t.requires_grad_(grad_req != "null")
        self._data = t
        self._grad = NDArray(torch.zeros_like(t, requires_grad=False))
        self._grad_req = grad_req
        from .. import autograd
        autograd._register_leaf(self)

    def backward(self, out_grad=None, retain_graph=False, train_mode=True):
        from .. import autograd
        autograd.backward([self], [out_grad] if out_grad is not None else None, retain_graph=retain_graph)

    def zero_grad(self):
        if self._grad is not None:
            self._grad._t.zero_()

    # ---- indexing -------------------------------------------------------------------------
    @staticmethod
    def _idx(key):
        if isinstance(key, NDArray):
            t = key._t
            return t.long() if not t.dtype == torch.bool else t
        if isinstance(key, tuple):
            return tuple(NDArray._idx(k) for k in key)
        return key

    def __getitem__(self, key):
        return NDArray(self._t[NDArray._idx(key)])

    def __setitem__(self, key, value):
        t = self._t
        tgt = t.detach() if t.requires_grad else t
        if isinstance(value, NDArray):
            v = value._t
            if v.device != tgt.device:
                v = v.to(tgt.device, non_blocking=True)
            if isinstance(key, slice) and key == slice(None):
                tgt.copy_(v)
            else:
                tgt[NDArray._idx(key)] = v.to(tgt.dtype)
        elif isinstance(value, (np.ndarray, list, tuple)):
            tgt[NDArray._idx(key)] = torch.as_tensor(np.asarray(value), dtype=tgt.dtype, device=tgt.device)
        else:
            tgt[NDArray._idx(key)] = value

    # ---- arithmetic -----------------------------------------------------------------------
    @staticmethod
    def _raw(x, like=None):
        if isinstance(x, NDArray):
            return x._t
        if isinstance(x, np.ndarray):
            return torch.as_tensor(x, device=like.device if like is not None else None)
        return x

    def _bin(self, other, fn, rev=False):
        a, b = self._t, NDArray._raw(other, self._data)
        if isinstance(b, torch.Tensor) and b.device != a.device:
            b = b.to(a.device)
        return NDArray(fn(b, a) if rev else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, lambda x, y: x - y, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, lambda x, y: x / y, True)
    def __mod__(self, o): return self._bin(o, torch.remainder)
    def __rmod__(self, o): return self._bin(o, lambda x, y: torch.remainder(torch.as_tensor(x, dtype=y.dtype, device=y.device) if not isinstance(x, torch.Tensor) else x, y), True)

    def __imod__(self, o):
        self._t.remainder_(o._t if isinstance(o, NDArray) else o)
        return self
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, lambda x, y: x ** y, True)
    def __neg__(self): return NDArray(-self._t)
    def __abs__(self): return NDArray(self._t.abs())
    def __matmul__(self, o): return self._bin(o, torch.matmul)

    def _cmp(self, o, fn):
        r = self._bin(o, fn)
        r._data = r._data.to(self._data.dtype if self._data.is_floating_point() else torch.float32)
        return r

    def __eq__(self, o): return self._cmp(o, torch.eq)
    def __ne__(self, o): return self._cmp(o, torch.ne)
    def __lt__(self, o): return self._cmp(o, torch.lt)
    def __le__(self, o): return self._cmp(o, torch.le)
    def __gt__(self, o): return self._cmp(o, torch.gt)
    def __ge__(self, o): return self._cmp(o, torch.ge)
    __hash__ = object.__hash__

    def _inplace(self, o, fn):
        t = self._t
        tgt = t.detach() if t.requires_grad else t
        b = NDArray._raw(o, tgt)
        if isinstance(b, torch.Tensor) and b.device != tgt.device:
            b = b.to(tgt.device)
        fn(tgt, b)
        return self

    def __iadd__(self, o): return self._inplace(o, lambda a, b: a.add_(b))
    def __isub__(self, o): return self._inplace(o, lambda a, b: a.sub_(b))
    def __imul__(self, o): return self._inplace(o, lambda a, b: a.mul_(b))
    def __itruediv__(self, o): return self._inplace(o, lambda a, b: a.div_(b))

    # ---- reductions & math ---------------------------------------------------------------
    def _red(self, fn, axis=None, keepdims=False):
        t = self._t
        if axis is None:
            return NDArray(fn(t).reshape(1))
        return NDArray(fn(t, dim=axis, keepdim=keepdims))

    def sum(self, axis=None, keepdims=False): return self._red(torch.sum, axis, keepdims)
    def mean(self, axis=None, keepdims=False): return self._red(torch.mean, axis, keepdims)

    def max(self, axis=None, keepdims=False):
        return NDArray(self._t.max().reshape(1)) if axis is None else NDArray(self._t.amax(dim=axis, keepdim=keepdims))

    def min(self, axis=None, keepdims=False):
        return NDArray(self._t.min().reshape(1)) if axis is None else NDArray(self._t.amin(dim=axis, keepdim=keepdims))

    def argmax(self, axis=None, keepdims=False):
        t = self._t
        r = t.reshape(-1).argmax().reshape(1) if axis is None else t.argmax(dim=axis, keepdim=keepdims)
        return NDArray(r.to(torch.float32))  # MXNet returns float indices

    def argmin(self, axis=None, keepdims=False):
        t = self._t
        r = t.reshape(-1).argmin().reshape(1) if axis is None else t.argmin(dim=axis, keepdim=keepdims)
        return NDArray(r.to(torch.float32))

    def norm(self, ord=2, axis=None, keepdims=False):
        t = self._t.float()
        return NDArray(torch.linalg.vector_norm(t, ord=ord).reshape(1) if axis is None
                       else torch.linalg.vector_norm(t, ord=ord, dim=axis, keepdim=keepdims))

    def abs(self): return NDArray(self._t.abs())
    def sqrt(self): return NDArray(self._t.sqrt())
    def square(self): return NDArray(self._t.square())
    def exp(self): return NDArray(self._t.exp())
    def log(self): return NDArray(self._t.log())
    def sign(self): return NDArray(self._t.sign())
    def relu(self): return NDArray(torch.relu(self._t))
    def sigmoid(self): return NDArray(torch.sigmoid(self._t))
    def tanh(self): return NDArray(torch.tanh(self._t))
    def clip(self, a_min, a_max): return NDArray(self._t.clamp(a_min, a_max))
    def softmax(self, axis=-1): return NDArray(torch.softmax(self._t, dim=axis))
    def log_softmax(self, axis=-1): return NDArray(torch.log_softmax(self._t, dim=axis))
    def dot(self, o): return dot(self, o)
    def one_hot(self, depth): return one_hot(self, depth)

    def topk(self, axis=-1, k=1, ret_typ="indices", is_ascend=False):
        return topk(self, axis=axis, k=k, ret_typ=ret_typ, is_ascend=is_ascend)


# ------------------------------------------------------------------------------------------
# creation
# ------------------------------------------------------------------------------------------
def from_torch(t: torch.Tensor) -> NDArray:
    return NDArray(t)


def array(source, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    ctx.check_available()
    if isinstance(source, NDArray):
        t = source._t.to(ctx.torch_device, copy=True)
        return NDArray(ctx.place(t if dtype is None else t.to(torch_dtype(dtype))), ctx)
    if isinstance(source, torch.Tensor):
        t = source.detach().to(ctx.torch_device, copy=True)          # mx.nd.array always copies (zero-copy wrapping is ``from_torch``)
        return NDArray(ctx.place(t if dtype is None else t.to(torch_dtype(dtype))), ctx)
    a = np.asarray(source)
    if dtype is None:
        dtype = a.dtype if isinstance(source, np.ndarray) and a.dtype != np.float64 else "float32"
    t = torch.tensor(a).to(torch_dtype(dtype))                       # torch.tensor copies: the NDArray never aliases the numpy buffer
    return NDArray(ctx.place(t.to(ctx.torch_device)), ctx)


def _shape(shape):
    return (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)


def zeros(shape, ctx=None, dtype=None, **kw):
    ctx = _ctx_of(ctx); ctx.check_available()
    return NDArray(ctx.place(torch.zeros(_shape(shape), dtype=torch_dtype(dtype), device=ctx.torch_device)), ctx)


def ones(shape, ctx=None, dtype=None, **kw):
    ctx = _ctx_of(ctx); ctx.check_available()
    return NDArray(ctx.place(torch.ones(_shape(shape), dtype=torch_dtype(dtype), device=ctx.torch_device)), ctx)


def empty(shape, ctx=None, dtype=None):
    ctx = _ctx_of(ctx); ctx.check_available()
    return NDArray(ctx.place(torch.empty(_shape(shape), dtype=torch_dtype(dtype), device=ctx.torch_device)), ctx)


def full(shape, val, ctx=None, dtype=None):
    ctx = _ctx_of(ctx); ctx.check_available()
    return NDArray(ctx.place(torch.full(_shape(shape), val, dtype=torch_dtype(dtype), device=ctx.torch_device)), ctx)


def arange(start, stop=None, step=1.0, repeat=1, ctx=None, dtype=None):
    ctx = _ctx_of(ctx); ctx.check_available()
    if stop is None:
        start, stop = 0, start
    t = torch.arange(start, stop, step, dtype=torch_dtype(dtype), device=ctx.torch_device)
    if repeat > 1:
        t = t.repeat_interleave(repeat)
    return NDArray(t, ctx)


def zeros_like(a): return NDArray(torch.zeros_like(a._t))
def ones_like(a): return NDArray(torch.ones_like(a._t))


def waitall():
    """Block until all queued work (KV scheduler + every CUDA stream) is complete.

    Parity: ``Engine::WaitForAll`` via ``mx.nd.waitall`` (``python/mxnet/ndarray/ndarray.py:156``)."""
    from ..kvstore import base as _kvb
    _kvb.flush_all()
    from .. import engine as _engine
    _engine.wait_all()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------
# functional ops (the subset of mx.nd.* user code and gluon blocks rely on)
# ------------------------------------------------------------------------------------------
def _w(t): return NDArray(t)


def concat(*arrays, dim=1): return _w(torch.cat([a._t for a in arrays], dim=dim))
def stack(*arrays, axis=0): return _w(torch.stack([a._t for a in arrays], dim=axis))


def dot(a, b, transpose_a=False, transpose_b=False):
    x, y = a._t, b._t
    if transpose_a: x = x.t()
    if transpose_b: y = y.t()
    return _w(torch.matmul(x, y))


def add_n(*arrays):
    out = arrays[0]._t.clone()
    for a in arrays[1:]:
        out.add_(a._t)
    return _w(out)


def elemwise_add(a, b): return a + b
def maximum(a, b): return _w(torch.maximum(a._t, NDArray._raw(b) if isinstance(b, NDArray) else torch.as_tensor(b, dtype=a._t.dtype, device=a._t.device)))
def minimum(a, b): return _w(torch.minimum(a._t, NDArray._raw(b) if isinstance(b, NDArray) else torch.as_tensor(b, dtype=a._t.dtype, device=a._t.device)))
def sqrt(a): return a.sqrt()
def square(a): return a.square()
def abs(a): return a.abs()
def exp(a): return a.exp()
def log(a): return a.log()
def sign(a): return a.sign()
def clip(a, a_min, a_max): return a.clip(a_min, a_max)
def where(c, a, b): return _w(torch.where(c._t != 0, a._t, b._t))
def softmax(a, axis=-1): return a.softmax(axis)
def log_softmax(a, axis=-1): return a.log_softmax(axis)
def relu(a): return a.relu()
def sigmoid(a): return a.sigmoid()
def tanh(a): return a.tanh()
def norm(a, ord=2, axis=None, keepdims=False): return a.norm(ord, axis, keepdims)
def moveaxis(a, s, d): return _w(torch.movedim(a._t, s, d))
def transpose(a, axes=None): return a.transpose(*(axes or ()))
def reshape(a, shape): return a.reshape(shape)
def flatten(a): return a.flatten()
def sum(a, axis=None, keepdims=False): return a.sum(axis, keepdims)
def mean(a, axis=None, keepdims=False): return a.mean(axis, keepdims)
def argmax(a, axis=None, keepdims=False): return a.argmax(axis, keepdims)
def argmin(a, axis=None, keepdims=False): return a.argmin(axis, keepdims)
def max(a, axis=None, keepdims=False): return a.max(axis, keepdims)
def min(a, axis=None, keepdims=False): return a.min(axis, keepdims)
def broadcast_to(a, shape): return a.broadcast_to(shape)
def tile(a, reps): return _w(a._t.repeat(*reps))
def cast(a, dtype): return a.astype(dtype)


def pick(data, index, axis=-1, keepdims=False):
    """``out[i] = data[i, index[i]]`` (``src/operator/tensor/broadcast_reduce_op_index.cc`` pick)."""
    idx = index._t.long().unsqueeze(axis)
    r = torch.gather(data._t, axis, idx)
    return _w(r if keepdims else r.squeeze(axis))


def one_hot(indices, depth, on_value=1.0, off_value=0.0, dtype="float32"):
    t = torch.nn.functional.one_hot(indices._t.long(), depth).to(torch_dtype(dtype))
    if on_value != 1.0 or off_value != 0.0:
        t = t * (on_value - off_value) + off_value
    return _w(t)


def topk(a, axis=-1, k=1, ret_typ="indices", is_ascend=False):
    v, i = torch.topk(a._t, k, dim=axis, largest=not is_ascend)
    if ret_typ == "value":
        return _w(v)
    if ret_typ == "both":
        return [_w(v), _w(i.float())]
    return _w(i.float())


def sort(a, axis=-1, is_ascend=True): return _w(torch.sort(a._t, dim=axis, descending=not is_ascend)[0])
def argsort(a, axis=-1, is_ascend=True): return _w(torch.sort(a._t, dim=axis, descending=not is_ascend)[1].float())
