"""Numerical test helpers (``mx.test_utils``).  Parity: ``python/mxnet/test_utils.py`` — ``assert_almost_equal``, ``almost_equal``,
``rand_ndarray``, ``rand_shape_nd``, ``numeric_grad`` / ``check_numeric_gradient`` (central differences against autograd),
``check_consistency`` (same computation on several contexts), ``default_context`` / ``set_default_context``."""
from __future__ import annotations

import numpy as np

from . import autograd, ndarray as nd
from .context import Context, cpu, current_context

__all__ = ["default_context", "set_default_context", "almost_equal", "assert_almost_equal", "rand_ndarray", "rand_shape_nd", "numeric_grad",
           "check_numeric_gradient", "check_consistency", "same"]

_default_ctx = None


def default_context():
    return _default_ctx or current_context()


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


def _np(a):
    return a.asnumpy() if hasattr(a, "asnumpy") else np.asarray(a)


def same(a, b):
    return np.array_equal(_np(a), _np(b))


def almost_equal(a, b, rtol=1e-5, atol=1e-20):
    return np.allclose(_np(a), _np(b), rtol=rtol, atol=atol)


def assert_almost_equal(a, b, rtol=1e-5, atol=1e-20, names=("a", "b")):
    a, b = _np(a), _np(b)
    if not np.allclose(a, b, rtol=rtol, atol=atol):
        err = np.abs(a - b) / (atol + rtol * np.abs(b))
        idx = np.unravel_index(np.argmax(err), err.shape) if err.shape else ()
        raise AssertionError("%s and %s differ: max violation %.3g at %s (%r vs %r), rtol=%g atol=%g" % (
            names[0], names[1], float(err.max()), idx, a[idx] if idx != () else a, b[idx] if idx != () else b, rtol, atol))


def rand_shape_nd(ndim, dim=10):
    return tuple(np.random.randint(1, dim + 1, size=ndim).tolist())


def rand_ndarray(shape, stype="default", density=None, dtype="float32", ctx=None):
    arr = nd.array(np.random.uniform(-1, 1, size=shape).astype(dtype), ctx=ctx or default_context())
    if stype == "row_sparse":
        keep = np.random.rand(shape[0]) < (0.5 if density is None else density)
        dense = arr.asnumpy(); dense[~keep] = 0
        return nd.array(dense, ctx=ctx or default_context()).tostype("row_sparse")
    return arr


def numeric_grad(f, inputs, eps=1e-3):
    """Central-difference gradient of the scalar ``f(*inputs)`` w.r.t. every numpy input."""
    grads = []
    for i, x in enumerate(inputs):
        g = np.zeros_like(x, dtype=np.float64)
        flat, gf = x.reshape(-1), g.reshape(-1)
        for j in range(flat.size):
            old = flat[j]
            flat[j] = old + eps; fp = float(f(*inputs))
            flat[j] = old - eps; fm = float(f(*inputs))
            flat[j] = old
            gf[j] = (fp - fm) / (2 * eps)
        grads.append(g)
    return grads


def check_numeric_gradient(fn, inputs, eps=1e-3, rtol=1e-2, atol=1e-3, ctx=None):
    """``fn(*NDArrays) -> NDArray``; compares autograd's gradient of ``fn(...).sum()`` with central differences (float64 accumulation)."""
    ctx = ctx or default_context()
    np_in = [np.array(_np(x), dtype=np.float32) for x in inputs]
    nds = [nd.array(x, ctx=ctx) for x in np_in]
    for x in nds:
        x.attach_grad()
    with autograd.record():
        out = fn(*nds).sum()
    out.backward()

    def f(*arrs):
        with autograd.pause():
            return fn(*[nd.array(a, ctx=ctx) for a in arrs]).sum().asscalar()
    for x, g in zip(nds, numeric_grad(f, np_in, eps)):
        assert_almost_equal(x.grad, g.astype(np.float32), rtol, atol, ("autograd", "numeric"))


def check_consistency(fn, inputs, ctx_list, rtol=1e-4, atol=1e-5):
    """Runs ``fn`` on every context of ``ctx_list`` with the same inputs and asserts that the outputs agree with the first one."""
    ref = None
    for ctx in ctx_list:
        ctx = ctx if isinstance(ctx, Context) else cpu()
        out = _np(fn(*[nd.array(_np(x), ctx=ctx) for x in inputs]))
        if ref is None:
            ref = out
        else:
            assert_almost_equal(out, ref, rtol, atol, (str(ctx), str(ctx_list[0])))
