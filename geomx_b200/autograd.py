"""``mx.autograd`` — record / pause / backward over torch's tape.

Parity: ``python/mxnet/autograd.py:122`` (record), ``:243`` (backward) and
``src/imperative/imperative.cc:191,278`` (RecordOp / Backward).  The reference builds
an nnvm gradient graph; here torch's C++ autograd engine is the tape and the
hand-written sm_100a kernels plug in as ``torch.autograd.Function`` nodes
(``geomx_b200/ops``).  ``grad_req='write'`` is honoured by rebinding the gradient
handle to the freshly produced tensor (no extra copy kernel); ``'add'`` accumulates.
"""
from __future__ import annotations

import threading
import weakref

import torch

__all__ = ["record", "pause", "train_mode", "predict_mode", "is_recording", "is_training", "backward",
           "mark_variables", "set_recording", "set_training", "grad"]

_st = threading.local()


def _state():
    if not hasattr(_st, "rec"):
        _st.rec, _st.train = False, False
    return _st


def is_recording():
    return _state().rec


def is_training():
    return _state().train


def set_recording(flag):
    s = _state(); old = s.rec; s.rec = bool(flag); return old


def set_training(flag):
    s = _state(); old = s.train; s.train = bool(flag); return old


class _Scope:
    def __init__(self, rec, train):
        self._rec, self._train = rec, train
        self._grad_ctx = None

    def __enter__(self):
        s = _state()
        self._old = (s.rec, s.train)
        if self._rec is not None:
            s.rec = self._rec
        if self._train is not None:
            s.train = self._train
        self._grad_ctx = torch.enable_grad() if s.rec else torch.no_grad()
        self._grad_ctx.__enter__()
        return self

    def __exit__(self, *a):
        self._grad_ctx.__exit__(*a)
        s = _state(); s.rec, s.train = self._old


def record(train_mode=True):
    return _Scope(True, train_mode)


def pause(train_mode=False):
    return _Scope(False, train_mode)


def train_mode():
    return _Scope(None, True)


def predict_mode():
    return _Scope(None, False)


# leaves with attached grads (weak) so that backward can honour grad_req
class _LeafRegistry:
    """Weak, identity-keyed set.  (A ``WeakSet`` compares the referents with ``==`` when the same array is registered twice, and NDArray's
    ``==`` is elementwise.)"""

    def __init__(self):
        self._refs = {}

    def add(self, nd):
        key = id(nd)
        if key not in self._refs:
            self._refs[key] = weakref.ref(nd, lambda _r, k=key: self._refs.pop(k, None))

    def __iter__(self):
        for r in list(self._refs.values()):
            obj = r()
            if obj is not None:
                yield obj


_leaves = _LeafRegistry()


def _register_leaf(nd):
    _leaves.add(nd)


def mark_variables(variables, gradients, grad_reqs="write"):
    from .ndarray import NDArray
    if isinstance(variables, NDArray):
        variables, gradients = [variables], [gradients]
    if isinstance(grad_reqs, str):
        grad_reqs = [grad_reqs] * len(variables)
    for v, g, r in zip(variables, gradients, grad_reqs):
        v.attach_grad(r)
        v._grad = g
        _register_leaf(v)


def backward(heads, head_grads=None, retain_graph=False, train_mode=True):
    """Run backward from ``heads``; fills ``x.grad`` of every array with an attached grad."""
    from .ndarray import NDArray
    if isinstance(heads, NDArray):
        heads = [heads]
    ts = [h._t for h in heads]
    gs = None
    if head_grads is not None:
        if isinstance(head_grads, NDArray):
            head_grads = [head_grads]
        gs = [None if g is None else g._t for g in head_grads]
    if gs is None:
        gs = [torch.ones_like(t) for t in ts]
    else:
        gs = [torch.ones_like(t) if g is None else g for t, g in zip(ts, gs)]
    live = [l for l in list(_leaves) if l._data.requires_grad]
    for l in live:
        l._data.grad = None
    torch.autograd.backward(ts, gs, retain_graph=retain_graph)
    for l in live:
        g = l._data.grad
        if g is None:
            continue
        if l._grad_req == "add" and l._grad is not None:
            l._grad._data.add_(g)
        elif l._grad is not None:
            if l._grad._data.data_ptr() != g.data_ptr():
                if getattr(l, "_stable_grad", False) or l._grad._data.shape != g.shape:
                    l._grad._data.copy_(g)
                else:
                    l._grad._data = g  # 'write': rebind, zero-copy
        l._data.grad = None


def grad(heads, variables, head_grads=None, retain_graph=None, create_graph=False, train_mode=True):
    from .ndarray import NDArray
    single = isinstance(variables, NDArray)
    if isinstance(heads, NDArray):
        heads = [heads]
    vs = [variables] if single else list(variables)
    ts = [h._t for h in heads]
    gs = [torch.ones_like(t) for t in ts] if head_grads is None else [g._t for g in (
        [head_grads] if isinstance(head_grads, NDArray) else head_grads)]
    out = torch.autograd.grad(ts, [v._t for v in vs], gs, retain_graph=retain_graph, create_graph=create_graph)
    res = [NDArray(o) for o in out]
    return res[0] if single else res


class Function:
    """User-defined differentiable function (parity: python/mxnet/autograd.py Function :365-500)::

        class sigmoid(mx.autograd.Function):
            def forward(self, x):
                y = 1 / (1 + mx.nd.exp(-x)); self.save_for_backward(y); return y
            def backward(self, dy):
                y, = self.saved_tensors; return dy * y * (1 - y)

    ``forward`` runs outside the tape on NDArrays; ``backward`` receives one gradient per output and returns one per input.  An instance
    can be called once per recorded use (it holds the saved tensors)."""

    def __init__(self):
        self._used = False
        self.saved_tensors = ()

    def save_for_backward(self, *args):
        self.saved_tensors = args

    def forward(self, *inputs):
        raise NotImplementedError

    def backward(self, *output_grads):
        raise NotImplementedError

    def __call__(self, *inputs):
        from .ndarray import NDArray
        assert not self._used, "Each Function instance can only be called once. Please create another instance."
        self._used = True
        user = self

        class _Bridge(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *ts):
                with pause():
                    outs = user.forward(*[NDArray(t.detach()) for t in ts])
                ctx.single = isinstance(outs, NDArray)
                outs = (outs,) if ctx.single else tuple(outs)
                res = tuple(o._t.detach() for o in outs)
                return res[0] if ctx.single else res

            @staticmethod
            def backward(ctx, *gs):
                with pause():
                    gi = user.backward(*[NDArray(g) for g in gs])
                gi = (gi,) if isinstance(gi, NDArray) else tuple(gi)
                assert len(gi) == len(inputs), "%s.backward must return exactly the same number of NDArrays as the number of NDArrays arguments to forward. " \
                    "Expecting %d got %d" % (type(user).__name__, len(inputs), len(gi))
                return tuple(None if g is None else g._t for g in gi)

        out = _Bridge.apply(*[x._t for x in inputs])
        return NDArray(out) if isinstance(out, torch.Tensor) else tuple(NDArray(o) for o in out)


def get_symbol(x):
    """The reference returns the recorded computation history as a Symbol; the tape here is PyTorch's and has no nnvm form."""
    from .base import MXNetError
    raise MXNetError("autograd.get_symbol is not available: the autograd tape is not an nnvm graph (build the model with mx.sym to get a Symbol)")
