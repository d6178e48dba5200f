"""``gluon.Block`` / ``HybridBlock`` / ``SymbolBlock``-free subset.

Parity: ``python/mxnet/gluon/block.py`` — name scoping (``_BlockScope`` :36-120, prefixes such
as ``sequential0_conv0_``), ``collect_params`` (regex select), child registration via
``__setattr__``, ``save_parameters`` / ``load_parameters`` with *structural* names
(``0.weight``; :315,356), legacy ``save_params``/``load_params``, ``initialize``, ``cast``,
``hybridize`` (here: arms the CUDA-graph step capture in ``geomx_b200.models.graphed``),
``summary``, deferred shape inference on first call.
"""
from __future__ import annotations

import re
import threading
from collections import OrderedDict


from .. import ndarray as nd
from ..ndarray import NDArray
from .parameter import DeferredInitializationError, Parameter, ParameterDict

__all__ = ["Block", "HybridBlock"]


class _BlockScope:
    _current = threading.local()
    _global_counter = {}

    def __init__(self, block):
        self._block = block
        self._counter = {}
        self._old_scope = None

    @staticmethod
    def create(prefix, params, hint):
        current = getattr(_BlockScope._current, "value", None)
        if current is None:
            if prefix is None:
                cnt = _BlockScope._global_counter.get(hint, 0)
                _BlockScope._global_counter[hint] = cnt + 1
                prefix = "%s%d_" % (hint, cnt)
            params = ParameterDict(prefix) if params is None else ParameterDict(params.prefix, params)
            return prefix, params
        if prefix is None:
            count = current._counter.get(hint, 0)
            prefix = "%s%d_" % (hint, count)
            current._counter[hint] = count + 1
        if params is None:
            parent = current._block.params
            params = ParameterDict(parent.prefix + prefix, parent._shared)
        else:
            params = ParameterDict(params.prefix, params)
        return current._block.prefix + prefix, params

    def __enter__(self):
        if self._block._empty_prefix:
            return self
        self._old_scope = getattr(_BlockScope._current, "value", None)
        _BlockScope._current.value = self
        return self

    def __exit__(self, *a):
        if self._block._empty_prefix:
            return
        _BlockScope._current.value = self._old_scope


from .. import profiler as _prof  # noqa: E402


class _HookHandle:
    def __init__(self, table, key):
        self._table, self._key = table, key

    def detach(self):
        self._table.pop(self._key, None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.detach()


class Block:
    def __init__(self, prefix=None, params=None):
        self._empty_prefix = prefix == ""
        self._prefix, self._params = _BlockScope.create(prefix, params, self._alias())
        self._name = self._prefix[:-1] if self._prefix.endswith("_") else self._prefix
        self._scope = _BlockScope(self)
        self._children = OrderedDict()
        self._reg_params = {}
        self._forward_hooks, self._forward_pre_hooks = OrderedDict(), OrderedDict()

    def _alias(self):
        return self.__class__.__name__.lower()

    def __repr__(self):
        mods = "\n".join("  (%s): %s" % (k, repr(v).replace("\n", "\n  ")) for k, v in self._children.items())
        return "%s(\n%s\n)" % (self.__class__.__name__, mods) if mods else "%s()" % self.__class__.__name__

    def __setattr__(self, name, value):
        if isinstance(value, Block):
            self.register_child(value, name)
        elif isinstance(value, Parameter):
            self._reg_params[name] = value
        object.__setattr__(self, name, value)

    @property
    def prefix(self): return self._prefix
    @property
    def name(self): return self._name
    @property
    def params(self): return self._params

    def name_scope(self):
        return self._scope

    def register_child(self, block, name=None):
        if name is None:
            name = str(len(self._children))
        self._children[name] = block

    def register_forward_hook(self, hook):
        """``hook(block, inputs, output)`` after every forward; returns a handle whose ``detach()`` removes it (gluon/block.py:350-380)."""
        self._forward_hooks[id(hook)] = hook
        return _HookHandle(self._forward_hooks, id(hook))

    def register_forward_pre_hook(self, hook):
        """``hook(block, inputs)`` before every forward; a tuple returned by the hook replaces the inputs."""
        self._forward_pre_hooks[id(hook)] = hook
        return _HookHandle(self._forward_pre_hooks, id(hook))

    def apply(self, fn):
        for c in self._children.values():
            c.apply(fn)
        fn(self)
        return self

    def collect_params(self, select=None):
        ret = ParameterDict(self._params.prefix)
        if not select:
            ret.update(self.params)
        else:
            pat = re.compile(select)
            ret.update({n: v for n, v in self.params.items() if pat.match(n)})
        for c in self._children.values():
            ret.update(c.collect_params(select=select))
        return ret

    def _collect_params_with_prefix(self, prefix=""):
        if prefix:
            prefix += "."
        ret = {prefix + k: v for k, v in self._reg_params.items()}
        for name, child in self._children.items():
            ret.update(child._collect_params_with_prefix(prefix + name))
        return ret

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        from .. import initializer
        self.collect_params().initialize(init if init is not None else initializer.Uniform(), ctx, verbose, force_reinit)

    def cast(self, dtype):
        for c in self._children.values():
            c.cast(dtype)
        for p in self.params.values():
            p.cast(dtype)

    def hybridize(self, active=True, **kwargs):
        for c in self._children.values():
            c.hybridize(active, **kwargs)

    def save_parameters(self, filename):
        params = self._collect_params_with_prefix()
        arg = {k: NDArray(v.data(v.list_ctx()[0])._t.detach()) for k, v in params.items() if v._data is not None}
        nd.save(filename, arg)

    def load_parameters(self, filename, ctx=None, allow_missing=False, ignore_extra=False):
        loaded = nd.load(filename)
        params = self._collect_params_with_prefix()
        if not loaded and not params:
            return
        if not any("." in k for k in loaded.keys()):
            # legacy format keyed by full parameter names
            self.collect_params().load(filename, ctx, allow_missing, ignore_extra, self.prefix)
            return
        if not allow_missing:
            for name in params:
                assert name in loaded, "Parameter '%s' is missing in file '%s'" % (name, filename)
        for name, v in loaded.items():
            if name not in params:
                if not ignore_extra:
                    raise ValueError("Parameter '%s' loaded from file '%s' is not present in this block" % (name, filename))
                continue
            p = params[name]
            if p._data is None and not p._deferred_init:
                p.shape = tuple(v.shape); p.initialize(ctx=ctx)
            p.set_data(v)

    def save_params(self, filename):
        self.collect_params().save(filename, strip_prefix=self.prefix)

    def load_params(self, filename, ctx=None, allow_missing=False, ignore_extra=False):
        self.load_parameters(filename, ctx, allow_missing, ignore_extra)

    def __call__(self, *args):
        for h in list(self._forward_pre_hooks.values()):
            r = h(self, args)
            if isinstance(r, tuple):
                args = r
        if _prof._state["running"] and (_prof._cfg["profile_imperative"] or _prof._cfg["profile_all"]):
            dev = bool(args) and hasattr(args[0], "_t") and args[0]._t.is_cuda
            with _prof.scope(self.name or type(self).__name__, "block", device=dev):
                out = self.forward(*args)
        else:
            out = self.forward(*args)
        for h in list(self._forward_hooks.values()):
            h(self, args, out)
        return out

    def forward(self, *args):
        raise NotImplementedError

    def summary(self, *inputs):
        total = 0
        print("%-40s %-20s %12s" % ("Layer (type)", "Shape", "Param #"))
        for name, p in self.collect_params().items():
            n = 1
            for s in (p.shape or ()):
                n *= s
            total += n
            print("%-40s %-20s %12d" % (name, str(p.shape), n))
        print("Total params: %d" % total)


class _FMod:
    """The ``F`` namespace handed to ``hybrid_forward`` (NDArray flavour)."""

    def __getattr__(self, name):
        return getattr(nd, name)


class _CachedOp:
    """One captured (forward graph, backward graph) pair of a HybridBlock for one input signature."""

    _tls = threading.local()

    @classmethod
    def inside(cls):
        return getattr(cls._tls, "depth", 0) > 0

    @classmethod
    def lookup(cls, block, args):
        import torch
        from .. import autograd
        if not args or not all(isinstance(a, NDArray) and a._t.is_cuda and a._t.dtype == torch.float32 for a in args) or _prof._state["running"]:
            return None
        try:
            params = [p for p in block.collect_params().values() if p._data is not None]
            if len(params) != len(block.collect_params()):
                return None                                   # deferred shapes: run eagerly once, capture on a later call
        except Exception:
            return None
        key = (tuple((tuple(a.shape), a._t.device.index) for a in args), autograd.is_recording(), autograd.is_training())
        op = block._cached_ops.get(key)
        if op is None:
            op = block._cached_ops[key] = cls(block, args, params, autograd.is_recording(), autograd.is_training())
        return op if op.ok else None

    def __init__(self, block, args, params, recording, training):
        import torch
        self.ok = False
        tensors = [t for p in params for t in [p.data(args[0].context)._t] if t.requires_grad]
        cached = self

        class Shim(torch.nn.Module):
            def parameters(self, recurse=True):            # the Gluon parameters are the graph's differentiable inputs
                return iter(tensors)

            def forward(self, *xs):
                cached._tls.depth = getattr(cached._tls, "depth", 0) + 1
                try:
                    out = Block.__call__(block, *[NDArray(x) for x in xs])
                finally:
                    cached._tls.depth -= 1
                self.multi = isinstance(out, (list, tuple))
                return tuple(o._t for o in out) if self.multi else out._t

        self.shim = Shim()
        from .. import autograd
        try:
            samples = tuple(a._t.detach().clone().requires_grad_(a._t.requires_grad) for a in args)
            with autograd._Scope(recording, training):
                self.fn = torch.cuda.make_graphed_callables(self.shim, samples, num_warmup_iters=3)
            self.ok = True
        except Exception as e:                                  # an op that cannot be captured: stay eager for this signature
            import warnings
            warnings.warn("hybridize(static_alloc=True): graph capture of %s failed (%s); running eagerly" % (block.name, e))

    def __call__(self, args):
        out = self.fn(*[a._t for a in args])
        return [NDArray(o) for o in out] if isinstance(out, tuple) else NDArray(out)


class HybridBlock(Block):
    """Block whose ``hybrid_forward(F, x, **params)`` receives its own parameters as kwargs."""

    def __init__(self, prefix=None, params=None):
        super().__init__(prefix, params)
        self._active = False
        self._static_alloc = False
        self._cached_ops = {}

    def hybridize(self, active=True, static_alloc=False, static_shape=False, **kwargs):
        """``hybridize()`` marks the block as traceable.  ``hybridize(static_alloc=True)`` is the CachedOp of this framework (reference:
        src/imperative/cached_op.cc — static memory planning + bulk execution of the cached graph): the block's forward AND backward are
        captured into CUDA graphs (``torch.cuda.make_graphed_callables``, one pair per input signature) with statically allocated
        activations, so a step replays two graph launches instead of one launch per operator.  Applies to the outermost block that was
        hybridized this way, on CUDA, for fixed input shapes; anything else falls back to eager execution."""
        self._active = active
        self._static_alloc = bool(active and static_alloc)
        self._cached_ops = {}
        super().hybridize(active, **kwargs)

    def __call__(self, *args):
        if self._static_alloc and not _CachedOp.inside():
            op = _CachedOp.lookup(self, args)
            if op is not None:
                return op(args)
        return super().__call__(*args)

    def infer_shape(self, *args):
        """Layers with deferred shapes override ``_infer(x)``."""
        self._infer(*args)

    def _infer(self, *args):
        pass

    def forward(self, x, *args):
        from ..symbol import Symbol
        if isinstance(x, Symbol):                     # symbolic tracing (export): build the graph instead of computing
            from . import _symbolic
            return _symbolic.call(self, x, *args)
        self._n_inputs = 1 + len(args)
        try:
            params = {k: v.data(x.context) if len(v.list_ctx()) > 1 else v.data() for k, v in self._reg_params.items()}
        except DeferredInitializationError:
            self._infer(x, *args)
            for p in self._reg_params.values():
                p._finish_deferred_init()
            params = {k: v.data() for k, v in self._reg_params.items()}
        return self.hybrid_forward(_F, x, *args, **params)

    def hybrid_forward(self, F, x, *args, **kwargs):
        raise NotImplementedError

    def infer_type(self, *args):
        """Parameters take the dtype of the first input (layers are dtype-generic; gluon/block.py infer_type)."""
        for p in self._reg_params.values():
            if p._data is None:
                p.dtype = str(args[0].dtype).replace("torch.", "") if args else p.dtype

    def export(self, path, epoch=0, nnvm=False):
        """``path-symbol.json`` + ``path-%04d.params`` (``arg:`` / ``aux:`` prefixed like ``save_checkpoint``) — loadable by
        ``SymbolBlock.imports``, ``mx.model.load_checkpoint``, ``Module`` and the C predict API.  A ``SymbolBlock`` writes the graph it
        owns; any other HybridBlock is traced symbolically (``gluon/_symbolic.py``) with inputs named ``data`` (``data0``, ``data1`` … when
        it was last called with several).  ``nnvm=True`` writes the JSON in the reference's dialect.  Parity: gluon/block.py:866-927."""
        from .. import symbol as S
        sym = getattr(self, "_symbol", None)
        if sym is None:
            n_in = getattr(self, "_n_inputs", 1)
            ins = [S.Variable("data")] if n_in == 1 else [S.Variable("data%d" % i) for i in range(n_in)]
            out = self(*ins)
            sym = S.Group(list(out)) if isinstance(out, (list, tuple)) else out
        with open("%s-symbol.json" % path, "w") as f:
            f.write(sym.tojson(nnvm=nnvm))
        aux, args = set(sym.list_auxiliary_states()), set(sym.list_arguments())
        d = {}
        for n, p in self.collect_params().items():
            if n in aux or n in args:
                d[("aux:" if n in aux else "arg:") + n] = p.data()
        nd.save("%s-%04d.params" % (path, epoch), d)
        return sym


_F = _FMod()


class SymbolBlock(HybridBlock):
    """A block built from a Symbol: ``SymbolBlock(outputs, inputs)`` turns every free argument / auxiliary state of ``outputs`` other than
    ``inputs`` into a Parameter (named as in the graph, no prefix) and evaluates the graph on call — differentiable, so exported models
    can be fine-tuned (parity: python/mxnet/gluon/block.py SymbolBlock :930-1100, incl. ``SymbolBlock.imports``)."""

    def __init__(self, outputs, inputs, params=None):
        super().__init__(prefix="", params=None)
        from .. import symbol as S
        self._symbol = S.Group(list(outputs)) if isinstance(outputs, (list, tuple)) and len(outputs) > 1 else (outputs[0] if isinstance(outputs, (list, tuple)) else outputs)
        ins = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        self._input_names = [i.name if isinstance(i, S.Symbol) else str(i) for i in ins]
        aux = set(self._symbol.list_auxiliary_states())
        self._param_names = [n for n in self._symbol.list_arguments() if n not in self._input_names]
        self._aux_names = list(self._symbol.list_auxiliary_states())
        for n in self._param_names + self._aux_names:
            p = (params.get(n) if params is not None and n in params else None) or Parameter(
                n, grad_req="null" if n in aux else "write", allow_deferred_init=True, differentiable=n not in aux)
            if n.endswith("_moving_var") or n.endswith("running_var"):
                from .. import initializer
                p.init = p.init or initializer.One()
            self._params._params[n] = p
            self._reg_params[n] = p

    @staticmethod
    def imports(symbol_file, input_names, param_file=None, ctx=None):
        from .. import symbol as S
        sym = S.load(symbol_file)
        names = [input_names] if isinstance(input_names, str) else list(input_names)
        blk = SymbolBlock(sym, [S.Variable(n) for n in names])
        if param_file is not None:
            loaded = nd.load(param_file)
            loaded = {k.split(":", 1)[1] if k.startswith(("arg:", "aux:")) else k: v for k, v in loaded.items()}
            for n, p in blk._reg_params.items():
                if n not in loaded:
                    raise RuntimeError("Parameter %s is missing in file %s" % (n, param_file))
                p.shape = tuple(loaded[n].shape)
                p.initialize(ctx=ctx)
                p.set_data(loaded[n])
        return blk

    def _shapes_from(self, args):
        shapes = {n: tuple(a.shape) for n, a in zip(self._input_names, args)}
        arg_shapes, _, aux_shapes = self._symbol.infer_shape(**shapes)
        known = dict(zip(self._symbol.list_arguments(), arg_shapes)); known.update(zip(self._aux_names, aux_shapes))
        for n, p in self._reg_params.items():
            if p.shape is None or 0 in tuple(p.shape) or p._data is None:
                p.shape = tuple(known[n])

    def forward(self, *args):
        from .. import autograd
        from ..symbol import run_graph
        assert len(args) == len(self._input_names), "expected %d inputs (%s)" % (len(self._input_names), self._input_names)
        try:
            pvals = {n: p.data(args[0].context) for n, p in self._reg_params.items()}
        except DeferredInitializationError:
            self._shapes_from(args)
            for p in self._reg_params.values():
                p._finish_deferred_init()
            pvals = {n: p.data(args[0].context) for n, p in self._reg_params.items()}
        feed = {n: a._t for n, a in zip(self._input_names, args)}
        feed.update({n: v._t for n, v in pvals.items()})
        outs = run_graph(self._symbol, feed, autograd.is_training())
        res = [NDArray(o) for o in outs]
        return res[0] if len(res) == 1 else res

    def hybrid_forward(self, F, x, *args, **kwargs):
        raise NotImplementedError
