"""``mx.operator`` — user-defined operators in Python (``CustomOp`` / ``CustomOpProp`` / ``register`` + ``mx.nd.Custom``).

Parity: ``python/mxnet/operator.py`` (CustomOp :421-470, CustomOpProp :472-640, register :642-780) over ``src/operator/custom/custom.cc``:
``forward(is_train, req, in_data, out_data, aux)`` / ``backward(req, out_grad, in_data, out_data, in_grad, aux)`` with ``self.assign(dst, req,
src)``, shapes from ``infer_shape``.  The reference runs the callbacks on a dedicated worker thread of its engine; here the operator is
wrapped in a ``torch.autograd.Function`` so it composes with the tape (``mx.autograd.record``) like any built-in op."""
from __future__ import annotations

import torch

from .base import MXNetError
from .ndarray import NDArray

__all__ = ["CustomOp", "CustomOpProp", "register", "get_all_registered_operators", "Custom", "PythonOp", "NumpyOp", "NDArrayOp"]

_registry = {}


class CustomOp:
    def forward(self, is_train, req, in_data, out_data, aux):
        raise NotImplementedError

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        raise NotImplementedError

    @staticmethod
    def assign(dst, req, src):
        if req in ("null", None):
            return
        s = src._t if isinstance(src, NDArray) else torch.as_tensor(src, dtype=dst._t.dtype, device=dst._t.device)
        if req in ("write", "inplace"):
            dst._t.copy_(s)
        elif req == "add":
            dst._t.add_(s)
        else:
            raise MXNetError("unknown req %s" % req)


class CustomOpProp:
    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad

    def list_arguments(self): return ["data"]
    def list_outputs(self): return ["output"]
    def list_auxiliary_states(self): return []
    def infer_shape(self, in_shape): return in_shape, [in_shape[0]] * len(self.list_outputs()), []
    def infer_type(self, in_type): return in_type, [in_type[0]] * len(self.list_outputs()), [in_type[0]] * len(self.list_auxiliary_states())
    def declare_backward_dependency(self, out_grad, in_data, out_data): return list(out_grad) + list(in_data) + list(out_data)
    def create_operator(self, ctx, in_shapes, in_dtypes): raise NotImplementedError


def register(reg_name):
    def deco(prop_cls):
        _registry[reg_name] = prop_cls
        return prop_cls
    return deco


def get_all_registered_operators():
    return sorted(_registry)


class _CustomFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, n_out, out_shapes, is_train, *ins):
        in_nd = [NDArray(t.detach()) for t in ins]
        outs = [NDArray(torch.zeros(tuple(s), dtype=ins[0].dtype, device=ins[0].device)) for s in out_shapes]
        op.forward(is_train, ["write"] * n_out, in_nd, outs, [])
        ctx.op, ctx.in_nd, ctx.outs = op, in_nd, outs
        res = tuple(o._t for o in outs)
        return res if n_out > 1 else res[0]

    @staticmethod
    def backward(ctx, *gout):
        in_grad = [NDArray(torch.zeros_like(i._t)) for i in ctx.in_nd]
        ctx.op.backward(["write"] * len(in_grad), [NDArray(g.contiguous()) for g in gout], ctx.in_nd, ctx.outs, in_grad, [])
        return (None, None, None, None) + tuple(g._t for g in in_grad)


def Custom(*inputs, op_type=None, **kwargs):
    """``mx.nd.Custom(x, ..., op_type='name', **params)``: instantiate the registered prop with the (string) params, infer the output shapes
    and run the operator; differentiable."""
    from . import autograd
    if op_type not in _registry:
        raise MXNetError("custom operator %r is not registered (have: %s)" % (op_type, ", ".join(get_all_registered_operators())))
    prop = _registry[op_type](**{k: str(v) for k, v in kwargs.items()})
    ins = [i._t for i in inputs]
    _, out_shapes, _ = prop.infer_shape([list(t.shape) for t in ins])
    op = prop.create_operator(inputs[0].context, [list(t.shape) for t in ins], [t.dtype for t in ins])
    n_out = len(prop.list_outputs())
    res = _CustomFn.apply(op, n_out, out_shapes, autograd.is_training(), *ins)
    return [NDArray(r) for r in res] if n_out > 1 else NDArray(res)


# ---- legacy operator classes (operator.py PythonOp :40-140, NumpyOp :143-250, NDArrayOp :253-400): the operator object itself carries
# forward / backward / infer_shape; calling it applies the op imperatively (differentiable), ``get_symbol`` gives the symbolic node.
class PythonOp:
    _seq = 0

    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad
        PythonOp._seq += 1
        self._reg_name = "_legacy_%s_%d" % (type(self).__name__, PythonOp._seq)
        outer = self

        class _Prop(CustomOpProp):
            def __init__(self):
                super().__init__(need_top_grad)

            def list_arguments(self): return outer.list_arguments()
            def list_outputs(self): return outer.list_outputs()
            def infer_shape(self, in_shape):
                r = outer.infer_shape(in_shape)
                return (r[0], r[1], []) if len(r) == 2 else r

            def create_operator(self, ctx, in_shapes, in_dtypes):
                return outer._make_op()
        _registry[self._reg_name] = _Prop

    def __call__(self, *args, **kwargs):
        return Custom(*args, op_type=self._reg_name)

    def get_symbol(self, *args, **kwargs):
        from . import symbol as sym
        name = kwargs.pop("name", None)
        return sym._nd_op("Custom")(*args, name=name, op_type=self._reg_name)

    def forward(self, in_data, out_data):
        out_data[0][:] = in_data[0]

    def backward(self, out_grad, in_data, out_data, in_grad):
        in_grad[0][:] = 1.0

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[0]]

    def list_outputs(self): return ["output"]
    def list_arguments(self): return ["data"]
    def need_top_grad(self): return self.need_top_grad_


class NumpyOp(PythonOp):
    """``forward(in_data, out_data)`` / ``backward(out_grad, in_data, out_data, in_grad)`` operate on numpy arrays (written in place)."""

    def _make_op(self):
        outer = self

        class _Op(CustomOp):
            def forward(self, is_train, req, in_data, out_data, aux):
                ins = [a.asnumpy() for a in in_data]; outs = [a.asnumpy().copy() for a in out_data]
                outer.forward(in_data=ins, out_data=outs)
                for d, r, o in zip(out_data, req, outs):
                    self.assign(d, r, o)

            def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                og = [a.asnumpy() for a in out_grad]; ins = [a.asnumpy() for a in in_data]; outs = [a.asnumpy() for a in out_data]
                ig = [a.asnumpy().copy() for a in in_grad]
                outer.backward(out_grad=og, in_data=ins, out_data=outs, in_grad=ig)
                for d, r, g in zip(in_grad, req, ig):
                    self.assign(d, r, g)
        return _Op()


class NDArrayOp(PythonOp):
    """Same protocol with NDArrays (the op may run on the GPU)."""

    def _make_op(self):
        outer = self

        class _Op(CustomOp):
            def forward(self, is_train, req, in_data, out_data, aux):
                outer.forward(in_data=in_data, out_data=out_data)

            def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                outer.backward(out_grad=out_grad, in_data=in_data, out_data=out_data, in_grad=in_grad)
        return _Op()

    def declare_backward_dependency(self, out_grad, in_data, out_data):
        deps = list(out_grad) if self.need_top_grad() else []
        return deps + list(in_data) + list(out_data)
