"""``mx.nd.random`` — uniform / normal / randint / shuffle / seed.

Parity: ``python/mxnet/ndarray/random.py`` (uniform, normal, randint, shuffle)."""
from __future__ import annotations

import torch

from .ndarray import NDArray, _ctx_of, _shape, torch_dtype

__all__ = ["uniform", "normal", "randn", "randint", "shuffle", "seed"]

_gens = {}


def _gen(dev):
    return _gens.get(str(dev))


def seed(seed_state, ctx="all"):
    torch.manual_seed(int(seed_state))
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(int(seed_state))


def uniform(low=0.0, high=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    ctx = _ctx_of(ctx); ctx.check_available()
    if out is not None:
        out._t.uniform_(low, high)
        return out
    t = torch.empty(_shape(shape), dtype=torch_dtype(dtype), device=ctx.torch_device).uniform_(low, high)
    return NDArray(t, ctx)


def normal(loc=0.0, scale=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    ctx = _ctx_of(ctx); ctx.check_available()
    if out is not None:
        out._t.normal_(loc, scale)
        return out
    t = torch.empty(_shape(shape), dtype=torch_dtype(dtype), device=ctx.torch_device).normal_(loc, scale)
    return NDArray(t, ctx)


def randn(*shape, **kw):
    return normal(kw.pop("loc", 0.0), kw.pop("scale", 1.0), shape=shape, **kw)


def randint(low, high, shape=(1,), dtype="int32", ctx=None, **kw):
    ctx = _ctx_of(ctx); ctx.check_available()
    return NDArray(torch.randint(int(low), int(high), _shape(shape), dtype=torch_dtype(dtype),
                                 device=ctx.torch_device), ctx)


def shuffle(data, **kw):
    perm = torch.randperm(data.shape[0], device=data._t.device)
    return NDArray(data._t[perm])
