"""Numerical test helpers (``mx.test_utils``).  Parity: ``python/mxnet/test_utils.py`` — ``assert_almost_equal``, ``almost_equal``,
``rand_ndarray``, ``rand_shape_nd``, ``numeric_grad`` / ``check_numeric_gradient`` (central differences against autograd),
``check_consistency`` (same computation on several contexts), ``default_context`` / ``set_default_context``."""
from __future__ import annotations

import numpy as np

from . import autograd, ndarray as nd
from .context import Context, cpu, current_context

__all__ = ["default_context", "set_default_context", "default_dtype", "get_atol", "get_rtol", "almost_equal", "assert_almost_equal",
           "almost_equal_ignore_nan", "assert_almost_equal_ignore_nan", "assert_exception", "find_max_violation", "same", "same_array",
           "rand_ndarray", "rand_sparse_ndarray", "create_sparse_array", "create_sparse_array_zd", "rand_shape_nd", "rand_shape_2d", "rand_shape_3d",
           "random_arrays", "random_sample", "np_reduce", "assign_each", "assign_each2", "shuffle_csr_column_indices", "numeric_grad",
           "check_numeric_gradient", "check_consistency", "simple_forward", "check_symbolic_forward", "check_symbolic_backward", "check_speed",
           "compare_optimizer", "compare_ndarray_tuple", "DummyIter", "list_gpus", "set_env_var", "EnvManager", "discard_stderr", "retry",
           "mean_check", "var_check", "chi_square_check", "gen_buckets_probs_with_ppf", "verify_generator", "download", "get_mnist",
           "get_mnist_ubyte", "get_mnist_iterator", "get_cifar10", "get_im2rec_path"]

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


# ------------------------------------------------------------------------------------------------ tolerances / comparisons
def default_dtype():
    return np.float32


def get_rtol(rtol=None):
    return 1e-5 if rtol is None else rtol


def get_atol(atol=None):
    return 1e-20 if atol is None else atol


def find_max_violation(a, b, rtol=None, atol=None):
    """Index and value of the largest ``|a-b| / (atol + rtol*|b|)``."""
    a, b = _np(a), _np(b)
    diff = np.abs(a - b)
    viol = diff / (get_atol(atol) + get_rtol(rtol) * np.abs(b) + 1e-20)
    idx = np.unravel_index(np.argmax(viol), viol.shape)
    return idx, float(np.max(viol))


def almost_equal_ignore_nan(a, b, rtol=None, atol=None):
    a, b = np.copy(_np(a)), np.copy(_np(b))
    nan = np.logical_or(np.isnan(a), np.isnan(b))
    a[nan] = 0; b[nan] = 0
    return almost_equal(a, b, get_rtol(rtol), get_atol(atol))


def assert_almost_equal_ignore_nan(a, b, rtol=None, atol=None, names=("a", "b")):
    a, b = np.copy(_np(a)), np.copy(_np(b))
    nan = np.logical_or(np.isnan(a), np.isnan(b))
    a[nan] = 0; b[nan] = 0
    assert_almost_equal(a, b, get_rtol(rtol), get_atol(atol), names)


def assert_exception(f, exception_type, *args, **kwargs):
    try:
        f(*args, **kwargs)
    except exception_type:
        return
    raise AssertionError("%s did not raise %s" % (getattr(f, "__name__", f), exception_type.__name__))


def same_array(array1, array2):
    """True when the two NDArrays share memory (a write through one is seen through the other)."""
    array1[:] = array1 + 1
    if not same(array1, array2):
        array1[:] = array1 - 1
        return False
    array1[:] = array1 - 1
    return same(array1, array2)


def compare_ndarray_tuple(t1, t2, rtol=None, atol=None):
    if t1 is None or t2 is None:
        return
    if isinstance(t1, tuple):
        for a, b in zip(t1, t2):
            compare_ndarray_tuple(a, b, rtol, atol)
    else:
        assert_almost_equal(t1, t2, get_rtol(rtol), get_atol(atol))


# ------------------------------------------------------------------------------------------------ random inputs
def rand_shape_2d(dim0=10, dim1=10):
    return np.random.randint(1, dim0 + 1), np.random.randint(1, dim1 + 1)


def rand_shape_3d(dim0=10, dim1=10, dim2=10):
    return np.random.randint(1, dim0 + 1), np.random.randint(1, dim1 + 1), np.random.randint(1, dim2 + 1)


def random_arrays(*shapes):
    arrs = [np.array(np.random.randn(), dtype=np.float32) if len(s) == 0 else np.random.randn(*s).astype(np.float32) for s in shapes]
    return arrs[0] if len(arrs) == 1 else arrs


def random_sample(population, k):
    population_copy = list(population)
    np.random.shuffle(population_copy)
    return population_copy[0:k]


def np_reduce(dat, axis, keepdims, numpy_reduce_func):
    """Apply a numpy reduction over one or several axes with ``keepdims`` handled uniformly."""
    axis = list(range(dat.ndim)) if axis is None else ([axis] if isinstance(axis, int) else list(axis))
    ret = dat
    for i in reversed(sorted(axis)):
        ret = numpy_reduce_func(ret, axis=i)
    if keepdims:
        shape = list(dat.shape)
        for i in axis:
            shape[i] = 1
        ret = ret.reshape(tuple(shape))
    return ret


def assign_each(the_input, function):
    return np.vectorize(function)(the_input) if function is not None else np.array(the_input)


def assign_each2(input1, input2, function):
    return np.vectorize(function)(input1, input2) if function is not None else np.array(input1)


def rand_sparse_ndarray(shape, stype, density=None, dtype=None, distribution=None, data_init=None, rsp_indices=None, modifier_func=None,
                        shuffle_csr_indices=False, ctx=None):
    """Random sparse array and its components: ``(array, (data, indices))`` for row_sparse, ``(array, (indptr, indices, data))`` for csr."""
    density = np.random.rand() if density is None else density
    dtype = dtype or np.float32
    ctx = ctx or default_context()
    if stype == "row_sparse":
        if rsp_indices is not None:
            idx = np.asarray(sorted(set(rsp_indices)), dtype=np.int64)
        else:
            idx = np.nonzero(np.random.rand(shape[0]) < density)[0].astype(np.int64)
        vals = np.random.uniform(-1, 1, size=(len(idx),) + tuple(shape[1:])).astype(dtype)
        if data_init is not None:
            vals[:] = data_init
        if modifier_func is not None:
            vals = assign_each(vals, modifier_func).astype(dtype)
        arr = nd.sparse.row_sparse_array((vals, idx), shape=shape, ctx=ctx, dtype=dtype)
        return arr, (vals, idx)
    if stype == "csr":
        assert len(shape) == 2
        import scipy.sparse as sp
        m = sp.random(shape[0], shape[1], density=density, format="csr", dtype=np.float64).astype(dtype)
        if data_init is not None:
            m.data[:] = data_init
        if modifier_func is not None:
            m.data = assign_each(m.data, modifier_func).astype(dtype)
        arr = nd.sparse.csr_matrix((m.data, m.indices, m.indptr), shape=shape, ctx=ctx, dtype=dtype)
        if shuffle_csr_indices:
            arr = shuffle_csr_column_indices(arr)
        return arr, (m.indptr, m.indices, m.data)
    raise ValueError("unknown storage type " + str(stype))


def create_sparse_array(shape, stype, data_init=None, rsp_indices=None, dtype=None, modifier_func=None, density=.5, shuffle_csr_indices=False):
    return rand_sparse_ndarray(shape, stype, density=density, data_init=data_init, rsp_indices=rsp_indices, dtype=dtype,
                               modifier_func=modifier_func, shuffle_csr_indices=shuffle_csr_indices)[0]


def create_sparse_array_zd(shape, stype, density, data_init=None, rsp_indices=None, dtype=None, modifier_func=None, shuffle_csr_indices=False):
    """Like ``create_sparse_array`` but tolerates ``density == 0`` (all-zero array)."""
    if stype == "row_sparse" and density == 0.0:
        rsp_indices = []
    return create_sparse_array(shape, stype, data_init, rsp_indices, dtype, modifier_func, density, shuffle_csr_indices)


def shuffle_csr_column_indices(csr):
    """Permute the column indices (and data) inside every row: same matrix, unsorted storage."""
    indptr, indices, data = csr.indptr.asnumpy().astype(np.int64), csr.indices.asnumpy().copy(), csr.data.asnumpy().copy()
    for r in range(len(indptr) - 1):
        lo, hi = indptr[r], indptr[r + 1]
        perm = np.random.permutation(hi - lo)
        indices[lo:hi] = indices[lo:hi][perm]; data[lo:hi] = data[lo:hi][perm]
    return nd.sparse.csr_matrix((data, indices, indptr), shape=csr.shape, dtype=data.dtype)


# ------------------------------------------------------------------------------------------------ symbolic checks
def _bind_location(sym, location, ctx, grad_req="null", aux_states=None):
    names = sym.list_arguments()
    if isinstance(location, (list, tuple)):
        location = dict(zip(names, location))
    args = {k: nd.array(_np(v), ctx=ctx) for k, v in location.items()}
    grads = {k: nd.zeros(v.shape, ctx=ctx) for k, v in args.items()} if grad_req != "null" else None
    aux = None
    if aux_states is not None:
        if isinstance(aux_states, (list, tuple)):
            aux_states = dict(zip(sym.list_auxiliary_states(), aux_states))
        aux = {k: nd.array(_np(v), ctx=ctx) for k, v in aux_states.items()}
    return sym.bind(ctx, args, args_grad=grads, grad_req=grad_req, aux_states=aux), args, grads


def simple_forward(sym, ctx=None, is_train=False, **inputs):
    """Forward a symbol on numpy inputs and return numpy output(s)."""
    ctx = ctx or default_context()
    ex = sym.bind(ctx, {k: nd.array(v, ctx=ctx) for k, v in inputs.items()})
    outs = [o.asnumpy() for o in ex.forward(is_train=is_train)]
    return outs[0] if len(outs) == 1 else outs


def check_symbolic_forward(sym, location, expected, rtol=1e-4, atol=None, aux_states=None, ctx=None, equal_nan=False, dtype=np.float32):
    """Bind ``sym`` at ``location`` (list or dict of numpy arrays), run forward and compare every output with ``expected``."""
    ctx = ctx or default_context()
    ex, _, _ = _bind_location(sym, location, ctx, "null", aux_states)
    outs = ex.forward(is_train=False)
    if isinstance(expected, dict):
        expected = [expected[k] for k in sym.list_outputs()]
    for name, exp, out in zip(sym.list_outputs(), expected, outs):
        (assert_almost_equal_ignore_nan if equal_nan else assert_almost_equal)(out, exp, rtol, get_atol(atol) if atol is None else atol, ("FORWARD_" + name, "EXPECTED_" + name))
    return [o.asnumpy() for o in outs]


def check_symbolic_backward(sym, location, out_grads, expected, rtol=1e-5, atol=None, aux_states=None, grad_req="write", ctx=None,
                            grad_stypes=None, equal_nan=False, dtype=np.float32):
    """Run forward(is_train) + backward(out_grads) and compare the argument gradients with ``expected`` (list or dict)."""
    ctx = ctx or default_context()
    ex, args, grads = _bind_location(sym, location, ctx, grad_req if isinstance(grad_req, str) else "write", aux_states)
    ex.forward(is_train=True)
    ogs = out_grads if isinstance(out_grads, (list, tuple)) else [out_grads]
    ex.backward([nd.array(_np(g), ctx=ctx) for g in ogs])
    if isinstance(expected, (list, tuple)):
        expected = dict(zip(sym.list_arguments(), expected))
    for name, exp in expected.items():
        (assert_almost_equal_ignore_nan if equal_nan else assert_almost_equal)(grads[name], exp, rtol, get_atol(atol) if atol is None else atol, ("BACKWARD_" + name, "EXPECTED_" + name))
    return {k: v.asnumpy() for k, v in grads.items()}


def check_speed(sym, location=None, ctx=None, N=20, grad_req=None, typ="whole", **kwargs):
    """Average seconds per ``forward`` (``typ='forward'``) or ``forward+backward`` (``'whole'``) of a bound symbol."""
    import time
    ctx = ctx or default_context()
    grad_req = grad_req or "write"
    if location is None:
        ex = sym.simple_bind(ctx, grad_req=grad_req, **kwargs)
        for a in ex.arg_arrays:
            a[:] = nd.random.normal(shape=a.shape)
    else:
        ex, _, _ = _bind_location(sym, location, ctx, grad_req)

    def run():
        if typ == "whole":
            ex.forward(is_train=True); ex.backward()
        else:
            ex.forward(is_train=False)
    run(); nd.waitall()
    tic = time.time()
    for _ in range(N):
        run()
    nd.waitall()
    return (time.time() - tic) / N


def compare_optimizer(opt1, opt2, shape, dtype, w_stype="default", g_stype="default", rtol=1e-4, atol=1e-5, compare_states=True):
    """One update of two optimizers from identical weights / gradients / states must agree."""
    w = np.random.uniform(-1, 1, size=shape).astype(dtype); g = np.random.uniform(-1, 1, size=shape).astype(dtype)
    w1, w2, g1, g2 = nd.array(w), nd.array(w), nd.array(g), nd.array(g)
    s1, s2 = opt1.create_state_multi_precision(0, w1), opt2.create_state_multi_precision(0, w2)
    opt1.update_multi_precision(0, w1, g1, s1); opt2.update_multi_precision(0, w2, g2, s2)
    if compare_states:
        compare_ndarray_tuple(s1 if isinstance(s1, tuple) else (s1,), s2 if isinstance(s2, tuple) else (s2,), rtol, atol)
    assert_almost_equal(w1, w2, rtol, atol)


# ------------------------------------------------------------------------------------------------ misc fixtures
class DummyIter:
    """Repeats the first batch of ``real_iter`` forever — for speed tests that should not measure IO."""

    def __init__(self, real_iter):
        self.real_iter = real_iter
        self.provide_data, self.provide_label, self.batch_size = real_iter.provide_data, real_iter.provide_label, real_iter.batch_size
        self.the_batch = next(iter(real_iter))

    def __iter__(self):
        return self

    def next(self):
        return self.the_batch

    __next__ = next

    def reset(self):
        pass


def list_gpus():
    import torch
    return list(range(torch.cuda.device_count())) if torch.cuda.is_available() else []


def set_env_var(key, val, default_val=""):
    import os
    prev = os.environ.get(key, default_val)
    os.environ[key] = val
    return prev


class EnvManager:
    """``with EnvManager('KEY', 'value'):`` — set an environment variable for the duration of the block."""

    def __init__(self, key, val):
        self._key, self._next, self._prev = key, val, None

    def __enter__(self):
        import os
        self._prev = os.environ.get(self._key)
        os.environ[self._key] = self._next

    def __exit__(self, *exc):
        import os
        if self._prev is not None:
            os.environ[self._key] = self._prev
        else:
            os.environ.pop(self._key, None)


class discard_stderr:
    """Silence file descriptor 2 inside the block (native code included)."""

    def __enter__(self):
        import os
        import sys
        sys.stderr.flush()
        self._saved = os.dup(2)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 2)

    def __exit__(self, *exc):
        import os
        import sys
        sys.stderr.flush()
        os.dup2(self._saved, 2)
        os.close(self._null); os.close(self._saved)


def retry(n):
    """Decorator: re-run a flaky (randomised) test up to ``n`` times, failing only if every attempt raises AssertionError."""
    assert n > 0

    def deco(orig):
        import functools

        @functools.wraps(orig)
        def wrapper(*args, **kwargs):
            for i in range(n):
                try:
                    return orig(*args, **kwargs)
                except AssertionError:
                    if i == n - 1:
                        raise
        return wrapper
    return deco


# ------------------------------------------------------------------------------------------------ statistical checks for samplers
def mean_check(generator, mu, sigma, nsamples=1000000):
    """Sample mean within 3 standard errors of ``mu``."""
    samples = np.array(generator(nsamples))
    return abs(samples.mean() - mu) < 3 * sigma / np.sqrt(nsamples)


def var_check(generator, sigma, nsamples=1000000):
    """Sample variance within 3 standard errors of ``sigma^2`` (normal approximation)."""
    samples = np.array(generator(nsamples))
    return abs(samples.var() - sigma ** 2) < 3 * np.sqrt(2 * sigma ** 4 / (nsamples - 1))


def gen_buckets_probs_with_ppf(ppf, nbuckets):
    """Equal-probability buckets of a continuous distribution from its percent point function."""
    probs = [1.0 / nbuckets] * nbuckets
    buckets = [(ppf(i / float(nbuckets)), ppf((i + 1) / float(nbuckets))) for i in range(nbuckets)]
    return buckets, probs


def chi_square_check(generator, buckets, probs, nsamples=1000000):
    """Pearson chi-square goodness of fit of ``generator(nsamples)`` against bucket probabilities.  Buckets are ``(lo, hi)`` intervals for
    continuous or plain values for discrete distributions.  Returns ``(p value, observed counts, expected counts)``."""
    import scipy.stats as ss
    samples = np.asarray(generator(nsamples)).reshape(-1)
    expected = np.asarray(probs, dtype=np.float64) * len(samples)
    if isinstance(buckets[0], (tuple, list)):
        obs = np.array([np.sum((samples >= lo) & (samples < hi)) for lo, hi in buckets], dtype=np.float64)
    else:
        obs = np.array([np.sum(samples == b) for b in buckets], dtype=np.float64)
    # samples outside every bucket are ignored on both sides
    expected = expected * obs.sum() / max(expected.sum(), 1e-30)
    _, p = ss.chisquare(f_obs=obs, f_exp=expected)
    return p, obs, expected


def verify_generator(generator, buckets, probs, nsamples=1000000, nrepeat=5, success_rate=0.15, alpha=0.05):
    """Repeat the chi-square check; at least ``success_rate`` of the runs must have p > alpha."""
    ps = [chi_square_check(generator, buckets, probs, nsamples)[0] for _ in range(nrepeat)]
    ok = sum(p > alpha for p in ps)
    if ok < nrepeat * success_rate:
        raise AssertionError("Generator test fails, Chi-square p=%s, buckets=%s, probs=%s" % (ps, buckets, probs))
    return ps


# ------------------------------------------------------------------------------------------------ data fixtures (never download)
def download(url, fname=None, dirname=None, overwrite=False, retries=5):
    """There is no network access in the environments this framework targets: returns the local file if it already exists, raises otherwise."""
    import os
    fname = fname or url.split("/")[-1]
    path = os.path.join(dirname, fname) if dirname else fname
    if os.path.exists(path):
        return path
    raise IOError("download(%s): no network access; place the file at %s" % (url, path))


def get_mnist(path=None):
    """``{'train_data', 'train_label', 'test_data', 'test_label'}`` as float32 NCHW in [0, 1] / int labels — real idx files under ``path``
    if present, otherwise the deterministic synthetic stand-in of ``gluon.data.vision.MNIST``."""
    from .gluon.data.vision import MNIST
    out = {}
    for split, train in (("train", True), ("test", False)):
        ds = MNIST(root=path or "data", train=train)
        out[split + "_data"] = ds._data.transpose(0, 3, 1, 2).astype(np.float32) / 255.0
        out[split + "_label"] = ds._label.astype(np.int64)
    return out


def get_mnist_ubyte(path="data"):
    import os
    need = ["train-images-idx3-ubyte", "train-labels-idx1-ubyte", "t10k-images-idx3-ubyte", "t10k-labels-idx1-ubyte"]
    missing = [f for f in need if not os.path.exists(os.path.join(path, f))]
    if missing:
        raise IOError("get_mnist_ubyte: %s missing under %s and there is no network access" % (missing, path))


def get_mnist_iterator(batch_size, input_shape, num_parts=1, part_index=0, path=None):
    """(train, val) NDArrayIters over this worker's ``part_index``-th of ``num_parts`` slices."""
    from . import io
    d = get_mnist(path)
    def shard(x):
        n = len(x) // num_parts
        return x[part_index * n:(part_index + 1) * n]
    tr = d["train_data"].reshape((-1,) + tuple(input_shape)); va = d["test_data"].reshape((-1,) + tuple(input_shape))
    return (io.NDArrayIter(shard(tr), shard(d["train_label"]).astype(np.float32), batch_size, shuffle=True),
            io.NDArrayIter(va, d["test_label"].astype(np.float32), batch_size))


def get_zip_data(data_dir, url, data_origin_name):
    """Extract ``data_origin_name`` (a .zip that must already be under ``data_dir``: nothing is downloaded) unless it has been extracted."""
    import os
    import zipfile
    os.makedirs(data_dir, exist_ok=True)
    path = download(url, fname=data_origin_name, dirname=data_dir)
    marker = os.path.join(data_dir, "." + data_origin_name + ".extracted")
    if not os.path.exists(marker):
        with zipfile.ZipFile(path) as z:
            z.extractall(data_dir)
        open(marker, "w").close()


def get_bz2_data(data_dir, data_name, url, data_origin_name):
    """Decompress ``data_origin_name`` (.bz2, already under ``data_dir``) to ``data_name`` unless that exists."""
    import bz2
    import os
    os.makedirs(data_dir, exist_ok=True)
    target = os.path.join(data_dir, data_name)
    if os.path.exists(target):
        return
    src = download(url, fname=data_origin_name, dirname=data_dir)
    with bz2.BZ2File(src) as fi, open(target, "wb") as fo:
        for chunk in iter(lambda: fi.read(1 << 20), b""):
            fo.write(chunk)


def get_mnist_pkl(path="data"):
    """The ``mnist.pkl.gz`` fixture of the reference's tests must be present locally (no network)."""
    import os
    if not os.path.exists(os.path.join(path, "mnist.pkl.gz")):
        raise IOError("get_mnist_pkl: %s/mnist.pkl.gz is missing and there is no network access; use get_mnist() for the synthetic stand-in" % path)


def get_cifar10(path="data"):
    from .gluon.data.vision import CIFAR10
    return CIFAR10(root=path, train=True), CIFAR10(root=path, train=False)


def get_im2rec_path(home_env="MXNET_HOME"):
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(here, "tools", "im2rec.py")
    if os.path.exists(p):
        return p
    raise IOError("tools/im2rec.py not found next to the package")
