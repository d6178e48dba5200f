"""``gluon.Parameter`` / ``ParameterDict``.

Parity: ``python/mxnet/gluon/parameter.py:418-548`` (data / grad / set_data / zero_grad /
list_data / list_grad / grad_req / initialize with deferred shape), ``ParameterDict``
(get, update, initialize, zero_grad, save/load, setattr)."""
from __future__ import annotations

from collections import OrderedDict

import torch

from .. import initializer as init_mod
from .. import ndarray as nd
from ..base import MXNetError
from ..context import Context, current_context
from ..ndarray import NDArray
from ..ndarray.ndarray import torch_dtype

tensor_types = None      # set below: the array types a Parameter accepts (python/mxnet/gluon/parameter.py:37)
__all__ = ["Parameter", "Constant", "ParameterDict", "DeferredInitializationError"]


class DeferredInitializationError(MXNetError):
    pass


class Parameter:
    def __init__(self, name, grad_req="write", shape=None, dtype="float32", lr_mult=1.0, wd_mult=1.0,
                 init=None, allow_deferred_init=False, differentiable=True, stype="default", grad_stype="default"):
        self.name = name
        self._grad_req = grad_req if differentiable else "null"
        self.shape = tuple(shape) if shape is not None else None
        self.dtype = dtype
        self.lr_mult, self.wd_mult = lr_mult, wd_mult
        self.init = init
        self.allow_deferred_init = allow_deferred_init
        self._differentiable = differentiable
        self._stype, self._grad_stype = stype, grad_stype
        self._data = None          # list[NDArray], one per ctx
        self._ctx_list = None
        self._deferred_init = ()
        self._trainer = None

    def __repr__(self):
        return "Parameter %s (shape=%s, dtype=%s)" % (self.name, self.shape, self.dtype)

    @property
    def grad_req(self):
        return self._grad_req

    @grad_req.setter
    def grad_req(self, req):
        assert req in ("write", "add", "null")
        if not self._differentiable:
            req = "null"
        if self._grad_req == req:
            return
        self._grad_req = req
        if self._data is not None:
            for d in self._data:
                if req == "null":
                    d._data = d._data.detach(); d._data.requires_grad_(False); d._grad = None; d._grad_req = "null"
                else:
                    d.attach_grad(req)

    def _shape_known(self):
        return self.shape is not None and all(s > 0 for s in self.shape)

    def initialize(self, init=None, ctx=None, default_init=None, force_reinit=False):
        if self._data is not None and not force_reinit:
            return
        default_init = default_init if default_init is not None else init_mod.Uniform()
        if ctx is None:
            ctx = [current_context()]
        if isinstance(ctx, Context):
            ctx = [ctx]
        if init is None:
            init = default_init if self.init is None else self.init
        if not self._shape_known():
            if self.allow_deferred_init:
                self._deferred_init = (init, ctx, default_init)
                return
            raise ValueError("Cannot initialize Parameter '%s' because it has invalid shape: %s." % (self.name, self.shape))
        self._deferred_init = (init, ctx, default_init)
        self._finish_deferred_init()

    def _finish_deferred_init(self):
        if not self._deferred_init:
            return
        init, ctx, default_init = self._deferred_init
        self._deferred_init = ()
        assert self._shape_known(), "Cannot initialize Parameter '%s' because it has invalid shape: %s." % (self.name, self.shape)
        host = torch.zeros(self.shape, dtype=torch.float32)
        initer = init_mod.create(init) if not callable(init) else init
        if isinstance(self.init, (init_mod.Initializer,)) and init is not self.init:
            initer = self.init
        initer(init_mod.InitDesc(self.name), host)
        self._init_impl(host, ctx)

    def _init_impl(self, host, ctx_list):
        self._ctx_list = list(ctx_list)
        self._data = []
        for c in self._ctx_list:
            c.check_available()
            t = host.to(torch_dtype(self.dtype)).to(c.torch_device).clone()
            a = NDArray(t, c)
            if self._grad_req != "null":
                a.attach_grad(self._grad_req)
            self._data.append(a)

    def _check_and_get(self, arr_list, ctx):
        if arr_list is not None:
            if ctx is list:
                return arr_list
            if ctx is None:
                if len(arr_list) == 1:
                    return arr_list[0]
                ctx = current_context()
            for c, a in zip(self._ctx_list, arr_list):
                if c == ctx:
                    return a
            raise RuntimeError("Parameter '%s' was not initialized on context %s. It was only initialized on %s."
                               % (self.name, str(ctx), str(self._ctx_list)))
        if self._deferred_init:
            raise DeferredInitializationError(
                "Parameter '%s' has not been initialized yet because initialization was deferred." % self.name)
        raise RuntimeError("Parameter '%s' has not been initialized." % self.name)

    def data(self, ctx=None):
        return self._check_and_get(self._data, ctx)

    def row_sparse_data(self, row_id):
        """The rows ``row_id`` of a ``row_sparse`` parameter as a RowSparseNDArray on ``row_id``'s context (parameter.py:480-500; with a
        kvstore-backed Trainer the rows are pulled from the server first)."""
        from ..ndarray.sparse import RowSparseNDArray
        from ..ndarray import NDArray
        import torch
        if self._stype != "row_sparse":
            raise RuntimeError("Cannot return a copy of Parameter %s via row_sparse_data() because its storage type is %s. Please use data() instead." % (self.name, self._stype))
        full = self._check_and_get(self._data, row_id.context)
        tr = getattr(self, "_trainer", None)
        if tr is not None and getattr(tr, "_row_sparse_pull", None) is not None:
            tr._row_sparse_pull(self, full, row_id)
        ids = torch.unique(row_id._t.long().reshape(-1))
        return RowSparseNDArray(NDArray(full._t.detach()[ids]), NDArray(ids), tuple(full.shape))

    def list_row_sparse_data(self, row_id):
        return [self.row_sparse_data(row_id)]

    def list_data(self):
        return list(self._check_and_get(self._data, list))

    def grad(self, ctx=None):
        d = self._check_and_get(self._data, ctx)
        if d._grad is None:
            raise RuntimeError("Cannot get gradient array for Parameter '%s' because grad_req='null'" % self.name)
        return d._grad

    def list_grad(self):
        ds = self._check_and_get(self._data, list)
        if ds[0]._grad is None:
            raise RuntimeError("Cannot get gradient array for Parameter '%s' because grad_req='null'" % self.name)
        return [d._grad for d in ds]

    def list_ctx(self):
        if self._data is None:
            if self._deferred_init:
                return self._deferred_init[1]
            raise RuntimeError("Parameter '%s' has not been initialized" % self.name)
        return self._ctx_list

    def set_data(self, data):
        self.shape = tuple(data.shape)
        if self._data is None:
            assert self._deferred_init, "Parameter '%s' has not been initialized" % self.name
            init, ctx, default_init = self._deferred_init
            self._deferred_init = ()
            self._init_impl(data._t.detach().float().cpu(), ctx)
            return
        for d in self._data:
            src = data._t.detach()
            if d._data.shape != src.shape or d._data.dtype != src.dtype:
                new = src.to(d._data.device).clone()
                d._data = new
                if self._grad_req != "null":
                    d.attach_grad(self._grad_req)
            else:
                d._data.detach().copy_(src)

    def zero_grad(self):
        if self._data is None:
            return
        for d in self._data:
            if d._grad is not None:
                d._grad._t.zero_()

    def reset_ctx(self, ctx):
        if isinstance(ctx, Context):
            ctx = [ctx]
        if self._data is not None:
            host = self._data[0]._t.detach().float().cpu()
            self._init_impl(host, ctx)
        elif self._deferred_init:
            init, _, default_init = self._deferred_init
            self._deferred_init = (init, ctx, default_init)

    def cast(self, dtype):
        self.dtype = dtype
        if self._data is None:
            return
        for d in self._data:
            d._data = d._data.detach().to(torch_dtype(dtype))
            if self._grad_req != "null":
                d.attach_grad(self._grad_req)

    def var(self):
        return self.name


class Constant(Parameter):
    def __init__(self, name, value):
        if not isinstance(value, NDArray):
            value = nd.array(value)
        self.value = value
        super().__init__(name, grad_req="null", shape=value.shape, dtype=value.dtype,
                         init=init_mod.Constant(value), differentiable=False)


class ParameterDict:
    def __init__(self, prefix="", shared=None):
        self._prefix = prefix
        self._params = OrderedDict()
        self._shared = shared

    def __repr__(self):
        return "%s(\n%s\n)" % (self._prefix, "\n".join("  " + repr(v) for v in self.values()))

    def __getitem__(self, key): return self._params[key]
    def __iter__(self): return iter(self._params)
    def __len__(self): return len(self._params)
    def __contains__(self, k): return k in self._params
    def items(self): return self._params.items()
    def keys(self): return self._params.keys()
    def values(self): return self._params.values()

    @property
    def prefix(self):
        return self._prefix

    def _get_impl(self, name):
        if name in self._params:
            return self._params[name]
        if self._shared is not None and name in self._shared._params:
            self._params[name] = self._shared._params[name]
            return self._shared._params[name]
        return None

    def get(self, name, **kwargs):
        name = self._prefix + name
        p = self._get_impl(name)
        if p is None:
            p = Parameter(name, **kwargs)
            self._params[name] = p
        else:
            for k, v in kwargs.items():
                if k == "shape" and v is not None and p.shape is not None:
                    # merge unknown dims
                    if len(v) == len(p.shape):
                        p.shape = tuple(a if a > 0 else b for a, b in zip(p.shape, v))
                elif getattr(p, k, None) is None:
                    setattr(p, k, v)
        return p

    def get_constant(self, name, value=None):
        name = self._prefix + name
        p = self._get_impl(name)
        if p is None:
            if value is None:
                raise KeyError("No constant named '%s'." % name)
            p = Constant(name, value); self._params[name] = p
        return p

    def update(self, other):
        for k, v in other.items():
            if k in self._params:
                assert self._params[k] is v, "Cannot update self with other because they have different Parameters with the same name '%s'" % k
            else:
                self._params[k] = v

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        init = init if init is not None else init_mod.Uniform()
        for v in self.values():
            v.initialize(None, ctx, init, force_reinit=force_reinit)

    def zero_grad(self):
        for v in self.values():
            v.zero_grad()

    def reset_ctx(self, ctx):
        for v in self.values():
            v.reset_ctx(ctx)

    def setattr(self, name, value):
        for v in self.values():
            setattr(v, name, value)

    def save(self, filename, strip_prefix=""):
        arg = {}
        for p in self.values():
            w = p.data(p.list_ctx()[0]) if p._data is not None else None
            if w is None:
                continue
            if not p.name.startswith(strip_prefix):
                raise ValueError("Prefix '%s' is to be striped before saving, but Parameter's name '%s' does not start with it" % (strip_prefix, p.name))
            arg[p.name[len(strip_prefix):]] = NDArray(w._t.detach(), w.context)
        nd.save(filename, arg)

    def load(self, filename, ctx=None, allow_missing=False, ignore_extra=False, restore_prefix=""):
        loaded = nd.load(filename)
        loaded = {restore_prefix + (k[4:] if k.startswith(("arg:", "aux:")) else k): v for k, v in loaded.items()}
        if not allow_missing:
            for name in self.keys():
                assert name in loaded, "Parameter '%s' is missing in file '%s'" % (name, filename)
        for name, v in loaded.items():
            if name not in self._params:
                assert ignore_extra, "Parameter '%s' loaded from file '%s' is not present in ParameterDict" % (name, filename)
                continue
            p = self._params[name]
            if p._data is None and not p._deferred_init:
                p.shape = tuple(v.shape)
                p.initialize(ctx=ctx)
            p.set_data(v)


from ..ndarray import NDArray as _NDArray  # noqa: E402
from ..symbol import Symbol as _Symbol  # noqa: E402
tensor_types = (_Symbol, _NDArray)
