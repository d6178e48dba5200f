"""Symbolic tracing of HybridBlocks: calling a block on ``mx.sym`` Symbols builds the graph that ``HybridBlock.export`` writes.

The reference hybridizes by running ``hybrid_forward(F=mx.symbol, ...)`` with ``Parameter.var()`` placeholders (python/mxnet/gluon/block.py
``_build_cache`` / ``export`` :860-927).  Here the built-in layers compute with kernels directly, so each of them has a small rule that emits
the equivalent graph node(s); user blocks written against ``F`` are traced exactly like the reference does.  Parameter variables carry the
parameters' own names, so the exported ``.params`` (``arg:`` / ``aux:`` + name) binds without translation."""
from ..base import MXNetError


def _rules():
    from .. import symbol as S
    from . import nn
    from .nn import basic_layers as B, conv_layers as C

    def act(y, kind, block):
        return y if kind is None else S.Activation(y, kind, name=block.name + "_" + kind)

    def dense(b, x, weight, bias=None):
        y = S.FullyConnected(x, b._units, weight=weight, bias=bias, no_bias=bias is None, flatten=b._flatten, name=b.name + "_fwd")
        return act(y, b._act, b)

    def conv2d(b, x, weight, bias=None):
        y = S.Convolution(x, b._kernel, b._channels, stride=b._strides, pad=b._padding, dilate=b._dilation, num_group=b._groups, weight=weight, bias=bias,
                          no_bias=bias is None, name=b.name + "_fwd")
        return act(y, b._act, b)

    def batchnorm(b, x, gamma, beta, running_mean, running_var):
        return S.Symbol("BatchNorm", b.name + "_fwd", [x, gamma, beta],
                        {"eps": float(b._eps), "momentum": float(b._momentum), "fix_gamma": False, "use_global_stats": bool(b._use_global), "axis": int(b._axis)},
                        [running_mean, running_var])

    def pool(kind):
        def rule(b, x):
            if getattr(b, "_ceil", False):
                return S._nd_op("Pooling")(x, kernel=b._k, pool_type=kind, stride=b._s, pad=b._p, pooling_convention="full", name=b.name + "_fwd")
            if kind == "avg" and not getattr(b, "_cip", True):
                return S._nd_op("Pooling")(x, kernel=b._k, pool_type=kind, stride=b._s, pad=b._p, count_include_pad=False, name=b.name + "_fwd")
            return S.Pooling(x, b._k, kind, b._s, b._p, name=b.name + "_fwd")
        return rule

    def gpool(kind):
        return lambda b, x: S.Pooling(x, (1, 1), kind, (1, 1), (0, 0), global_pool=True, name=b.name + "_fwd")

    rules = {
        B.Dense: dense,
        C.Conv2D: conv2d,
        B.BatchNorm: batchnorm,
        B.Activation: lambda b, x: S.Activation(x, b._act_type, name=b.name + "_fwd"),
        B.Dropout: lambda b, x: S.Dropout(x, b._rate, name=b.name + "_fwd"),
        B.Flatten: lambda b, x: S.Flatten(x, name=b.name + "_fwd"),
        B.LeakyReLU: lambda b, x: S._nd_op("LeakyReLU")(x, act_type="leaky", slope=float(b._alpha), name=b.name + "_fwd"),
        B.ELU: lambda b, x: S._nd_op("LeakyReLU")(x, act_type="elu", slope=float(b._alpha), name=b.name + "_fwd"),
        B.Embedding: lambda b, x, weight: S._nd_op("Embedding")(x, weight, name=b.name + "_fwd"),
        B.LayerNorm: lambda b, x, gamma, beta: S._nd_op("LayerNorm")(x, gamma, beta, axis=b._axis, eps=b._eps, name=b.name + "_fwd"),
        C.MaxPool2D: pool("max"),
        C.AvgPool2D: pool("avg"),
        C.GlobalAvgPool2D: gpool("avg"),
        C.GlobalMaxPool2D: gpool("max"),
    }
    for cls_name, kind in (("GlobalAvgPool2D", "avg"), ("GlobalMaxPool2D", "max")):
        cls = getattr(nn, cls_name, None)
        if cls is not None:
            rules.setdefault(cls, gpool(kind))
    return rules


_RULES = None


def call(block, x, *args):
    """``block(x_symbol, ...)``: the graph of this block applied to Symbols."""
    global _RULES
    from .. import symbol as S
    if _RULES is None:
        _RULES = _rules()
    params = {k: S.Variable(p.name) for k, p in block._reg_params.items()}
    rule = None
    for cls in type(block).__mro__:
        if cls in _RULES:
            # a subclass that overrides hybrid_forward is a different computation: trace it instead of applying the parent's rule
            if type(block).hybrid_forward is cls.hybrid_forward:
                rule = _RULES[cls]
            break
    try:
        if rule is not None:
            return rule(block, x, *args, **params)
        return block.hybrid_forward(S, x, *args, **params)
    except MXNetError:
        raise
    except (AttributeError, TypeError) as e:
        if "_t" in str(e) or "Symbol" in str(e):
            raise MXNetError("%s (%s) computes with tensors directly and has no symbolic rule; it cannot be part of an exported graph"
                             % (block.name, type(block).__name__)) from e
        raise
