"""``mx.nd._internal`` — the underscore-prefixed operators the reference's front end calls directly (``python/mxnet/ndarray/_internal.py``):
scalar arithmetic (``_plus_scalar`` ...), comparisons, creation (``_zeros`` / ``_ones`` / ``_full`` / ``_arange``), ``_copyto``, ``_set_value``
and the ``_random_*`` / ``_sample_*`` samplers.  Resolved lazily onto the public implementations."""
import torch

from .ndarray import NDArray

_SCALAR = {
    "_plus_scalar": lambda t, s: t + s, "_minus_scalar": lambda t, s: t - s, "_rminus_scalar": lambda t, s: s - t, "_mul_scalar": lambda t, s: t * s,
    "_div_scalar": lambda t, s: t / s, "_rdiv_scalar": lambda t, s: s / t, "_mod_scalar": lambda t, s: torch.remainder(t, s),
    "_rmod_scalar": lambda t, s: torch.remainder(torch.full_like(t, s), t), "_power_scalar": lambda t, s: t ** s, "_rpower_scalar": lambda t, s: s ** t,
    "_maximum_scalar": lambda t, s: torch.clamp(t, min=s), "_minimum_scalar": lambda t, s: torch.clamp(t, max=s),
    "_hypot_scalar": lambda t, s: torch.hypot(t, torch.full_like(t, s)),
    "_equal_scalar": lambda t, s: (t == s).to(t.dtype), "_not_equal_scalar": lambda t, s: (t != s).to(t.dtype),
    "_greater_scalar": lambda t, s: (t > s).to(t.dtype), "_greater_equal_scalar": lambda t, s: (t >= s).to(t.dtype),
    "_lesser_scalar": lambda t, s: (t < s).to(t.dtype), "_lesser_equal_scalar": lambda t, s: (t <= s).to(t.dtype),
    "_logical_and_scalar": lambda t, s: ((t != 0) & bool(s)).to(t.dtype), "_logical_or_scalar": lambda t, s: ((t != 0) | bool(s)).to(t.dtype),
    "_logical_xor_scalar": lambda t, s: ((t != 0) ^ bool(s)).to(t.dtype),
}
_BINARY = {"_plus": "add", "_minus": "sub", "_mul": "mul", "_div": "div", "_mod": "remainder", "_power": "pow", "_maximum": "maximum", "_minimum": "minimum",
           "_hypot": "hypot"}
_COMPARE = {"_equal": torch.eq, "_not_equal": torch.ne, "_greater": torch.gt, "_greater_equal": torch.ge, "_lesser": torch.lt, "_lesser_equal": torch.le}


def _out(res, out):
    if out is not None:
        out._t.copy_(res)
        return out
    return NDArray(res)


def __getattr__(name):
    from . import ndarray as nd, random as rnd
    if name in _SCALAR:
        fn = _SCALAR[name]
        return lambda data, scalar, out=None, **kw: _out(fn(data._t, float(scalar)), out)
    if name in _BINARY:
        fn = getattr(torch, _BINARY[name])
        return lambda lhs, rhs, out=None, **kw: _out(fn(lhs._t, rhs._t), out)
    if name in _COMPARE:
        fn = _COMPARE[name]
        return lambda lhs, rhs, out=None, **kw: _out(fn(lhs._t, rhs._t).to(lhs._t.dtype), out)
    if name in ("_zeros", "_ones", "_empty"):
        make = getattr(nd, name[1:])
        return lambda shape=(), ctx=None, dtype=None, out=None, **kw: make(shape, ctx=ctx, dtype=dtype) if out is None else _out(make(out.shape, ctx=out.context, dtype=out.dtype)._t, out)
    if name == "_full":
        return lambda shape, value, ctx=None, dtype=None, out=None, **kw: nd.full(shape, value, ctx=ctx, dtype=dtype) if out is None else _out(torch.full_like(out._t, value), out)
    if name == "_arange":
        return lambda start, stop=None, step=1.0, repeat=1, ctx=None, dtype=None, **kw: nd.arange(start, stop, step, repeat, ctx=ctx, dtype=dtype)
    if name == "_copyto":
        return lambda data, out=None, **kw: data.copyto(out) if out is not None else data.copy()
    if name == "_set_value":
        return lambda src, out=None, **kw: _out(torch.full_like(out._t, float(src)), out)
    if name in ("_copy", "_identity_with_attr_like_rhs"):
        return lambda data, *a, **kw: data.copy()
    for prefix in ("_random_", "_sample_"):
        if name.startswith(prefix) and hasattr(rnd, name[len(prefix):]):
            return getattr(rnd, name[len(prefix):])
    public = getattr(nd, name.lstrip("_"), None)          # e.g. _slice_assign helpers are not needed; plain aliases are
    if callable(public):
        return public
    raise AttributeError("mx.nd._internal has no operator %r" % name)
