"""``mx.nd.random`` — uniform / normal / randint / shuffle / seed + the sampling families (exponential, gamma, poisson,
negative_binomial, generalized_negative_binomial, multinomial).

Parity: ``python/mxnet/ndarray/random.py`` and ``src/operator/random/{sample_op,multisample_op,sample_multinomial_op}``: scalar parameters
draw ``shape`` samples; NDArray parameters draw ``shape`` samples PER parameter element (output ``params.shape + shape``)."""
from __future__ import annotations

import torch

from .ndarray import NDArray, _ctx_of, _shape, torch_dtype

__all__ = ["uniform", "normal", "randn", "randint", "shuffle", "seed", "exponential", "gamma", "poisson", "negative_binomial",
           "generalized_negative_binomial", "multinomial", "uniform_like", "normal_like", "exponential_like", "gamma_like", "poisson_like",
           "negative_binomial_like", "generalized_negative_binomial_like", "unique_zipfian"]

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


def _param_sampler(draw, params, shape, dtype, ctx):
    """Scalar params → ``shape`` samples on ``ctx``; NDArray params → ``params.shape + shape`` samples (multisample_op.h)."""
    shp = _shape(shape) if shape is not None else ()
    if any(isinstance(p, NDArray) for p in params):
        ts = [p._t.float() if isinstance(p, NDArray) else torch.tensor(float(p)) for p in params]
        ref = next(p for p in params if isinstance(p, NDArray))._t
        ts = [t.to(ref.device).expand(ref.shape) for t in ts]
        full = tuple(ref.shape) + (shp if shape is not None and shp != (1,) else ())
        ts = [t.reshape(tuple(ref.shape) + (1,) * (len(full) - ref.dim())).expand(full) for t in ts]
        return NDArray(draw(*ts).to(torch_dtype(dtype)))
    ctx = _ctx_of(ctx); ctx.check_available()
    ts = [torch.full(shp or (1,), float(p), device=ctx.torch_device) for p in params]
    return NDArray(draw(*ts).to(torch_dtype(dtype)), ctx)


def exponential(scale=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    """Exponential with mean ``scale`` (= 1/lambda)."""
    return _param_sampler(lambda s: torch.empty_like(s).exponential_(1.0) * s, [scale], shape, dtype, ctx)


def gamma(alpha=1.0, beta=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    """Gamma with shape ``alpha`` and SCALE ``beta``."""
    return _param_sampler(lambda a, b: torch.distributions.Gamma(a, 1.0 / b).sample(), [alpha, beta], shape, dtype, ctx)


def poisson(lam=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    return _param_sampler(lambda l: torch.poisson(l), [lam], shape, dtype, ctx)


def negative_binomial(k=1, p=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    """Failures before the ``k``-th success with success probability ``p``: Poisson(Gamma(k, (1-p)/p))."""
    return _param_sampler(lambda kk, pp: torch.poisson(torch.distributions.Gamma(kk, pp / (1 - pp).clamp_min(1e-12)).sample()), [k, p], shape, dtype, ctx)


def generalized_negative_binomial(mu=1.0, alpha=1.0, shape=(1,), dtype=None, ctx=None, out=None, **kw):
    """Mean ``mu``, dispersion ``alpha``: Poisson(Gamma(1/alpha, scale = alpha * mu)); ``alpha == 0`` degenerates to Poisson(mu)."""
    def draw(m, a):
        lam = torch.where(a > 0, torch.distributions.Gamma(1.0 / a.clamp_min(1e-12), 1.0 / (a * m).clamp_min(1e-12)).sample(), m)
        return torch.poisson(lam)
    return _param_sampler(draw, [mu, alpha], shape, dtype, ctx)


def multinomial(data, shape=None, get_prob=False, out=None, dtype="int32", **kw):
    """Sample category indices from the probability rows of ``data [..., K]``; ``shape`` samples per row.  ``get_prob`` additionally returns
    the log-probabilities of the draws (for REINFORCE-style estimators)."""
    p = data._t.float()
    n = 1
    shp = ()
    if shape is not None:
        shp = _shape(shape); n = 1
        for s in shp:
            n *= s
    idx = torch.multinomial(p.reshape(-1, p.shape[-1]), n, replacement=True)
    idx = idx.reshape(tuple(p.shape[:-1]) + shp)
    res = NDArray(idx.to(torch_dtype(dtype)))
    if get_prob:
        logp = torch.log(p.reshape(-1, p.shape[-1]).gather(1, idx.reshape(-1, n))).reshape(idx.shape)
        return res, NDArray(logp)
    return res


def _like(fn):
    def f(data, *args, **kwargs):
        kwargs.pop("shape", None)
        return fn(*args, shape=tuple(data.shape), dtype=str(data._t.dtype).replace("torch.", ""), ctx=data.context, **kwargs)
    f.__doc__ = "Samples with the shape / dtype / context of ``data`` (``_random_*_like``)."
    return f


uniform_like, normal_like, exponential_like, gamma_like = _like(uniform), _like(normal), _like(exponential), _like(gamma)
poisson_like, negative_binomial_like, generalized_negative_binomial_like = _like(poisson), _like(negative_binomial), _like(generalized_negative_binomial)


def unique_zipfian(range_max, shape=None, ctx=None, **kw):
    """``shape = (batch, n)`` rows of ``n`` DISTINCT classes drawn from an approximately Zipfian (log-uniform) distribution over
    ``[0, range_max)`` — the candidate sampler of sampled softmax (``_sample_unique_zipfian``).  Returns ``(samples, number of trials per row)``;
    the trial counts let the caller compute expected counts."""
    import math
    shp = _shape(shape)
    batch, n = (1, shp[0]) if len(shp) == 1 else shp
    assert n <= range_max, "cannot draw %d unique classes out of %d" % (n, range_max)
    out = torch.empty((batch, n), dtype=torch.int64); trials = torch.zeros(batch, dtype=torch.int64)
    log_range = math.log(range_max + 1)
    for b in range(batch):
        seen, k = {}, 0
        while len(seen) < n:
            draw = (torch.exp(torch.rand(2 * n) * log_range).long() - 1) % range_max
            for v in draw.tolist():
                k += 1
                if v not in seen:
                    seen[v] = len(seen)
                    if len(seen) == n:
                        break
        out[b] = torch.tensor(list(seen.keys())); trials[b] = k
    res = out if len(shp) == 2 else out[0]
    return NDArray(res), NDArray(trials)
