"""``mx.nd.sparse`` — row_sparse storage (the format ``kv.row_sparse_pull`` / sparse gradient pushes use).

Parity: ``python/mxnet/ndarray/sparse.py`` (``RowSparseNDArray`` :561-840, ``row_sparse_array`` :1040-1180, ``zeros``), the
``sparse_retain`` operator (``src/operator/tensor/sparse_retain-inl.h``) and ``cast_storage`` dense<->row_sparse
(``src/operator/tensor/cast_storage-inl.h``).  A RowSparseNDArray of logical shape ``(R, *row)`` stores ``data`` ``(nnz, *row)`` and sorted,
unique ``indices`` ``(nnz,)`` int64; rows that are not listed are zero.  CSR is represented through ``torch.sparse_csr`` on demand only
(no GeoMX path uses it)."""
from __future__ import annotations

import numpy as np
import torch

from ..base import MXNetError
from ..context import current_context
from .ndarray import NDArray, array as _dense_array


def _gather(src, ids):
    if src.is_cuda and src.dtype == torch.float32 and src.is_contiguous():
        from ..ops import native
        if native.available():
            from ..ops import _native_api
            return _native_api.gather_rows(src, ids)
    return src.index_select(0, ids.to(src.device))


class BaseSparseNDArray:
    """What the two sparse storage types share (python/mxnet/ndarray/sparse.py BaseSparseNDArray :100-250): size / ndim, dtype casts,
    copies to other arrays or contexts, format validation.  Element-wise maths densifies (``tostype('default')``) unless a sparse kernel exists."""

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def ndim(self):
        return len(self.shape)

    def astype(self, dtype, copy=True):
        return self._rebuild(self.data.astype(dtype))

    def as_in_context(self, ctx):
        return self._rebuild(self.data.as_in_context(ctx), move=ctx)

    def copyto(self, other):
        """Into another array of the same storage type (in place), into a dense NDArray, or onto a Context (a new sparse array there)."""
        from ..context import Context
        if isinstance(other, Context):
            return self.as_in_context(other)
        if isinstance(other, NDArray):
            other._t.copy_(self.tostype("default")._t)
            return other
        if type(other) is not type(self):
            raise TypeError("copyto does not support type %s" % type(other))
        fresh = self.copy()
        other.__dict__.update(fresh.__dict__)
        return other

    def reshape(self, *a, **k):
        raise NotImplementedError("reshape is not supported for sparse arrays: convert with tostype('default') first")

    def wait_to_read(self):
        self.data.wait_to_read()

    def check_format(self, full_check=True):
        """Raise if the auxiliary arrays do not describe a valid array of this storage type (sorted, in range, consistent lengths)."""
        self._check(full_check)


class RowSparseNDArray(BaseSparseNDArray):
    stype = "row_sparse"

    def __init__(self, data: NDArray, indices: NDArray, shape):
        self.data, self.indices, self._shape = data, indices, tuple(int(s) for s in shape)
        if data.shape[0] != indices.shape[0]:
            raise MXNetError("row_sparse: data has %d rows but %d indices" % (data.shape[0], indices.shape[0]))

    # ---- introspection
    @property
    def shape(self): return self._shape
    @property
    def dtype(self): return self.data.dtype
    @property
    def context(self): return self.data.context
    ctx = context
    @property
    def size(self): return int(np.prod(self._shape))
    def __repr__(self):
        return "<RowSparseNDArray %s @%s nnz_rows=%d>" % ("x".join(map(str, self._shape)), self.context, self.indices.shape[0])

    # ---- conversion
    def tostype(self, stype):
        if stype == "row_sparse":
            return self
        if stype != "default":
            raise MXNetError("cast_storage row_sparse -> %s is not supported" % stype)
        out = torch.zeros(self._shape, dtype=self.data._t.dtype, device=self.data._t.device)
        if self.indices.shape[0]:
            out.index_copy_(0, self.indices._t.long(), self.data._t)
        return NDArray(out)

    todense = lambda self: self.tostype("default")
    def asnumpy(self): return self.tostype("default").asnumpy()
    def wait_to_read(self): self.data.wait_to_read()
    def copy(self): return RowSparseNDArray(self.data.copy(), self.indices.copy(), self._shape)
    def as_in_context(self, ctx): return RowSparseNDArray(self.data.as_in_context(ctx), self.indices.as_in_context(ctx), self._shape)
    def astype(self, dtype): return RowSparseNDArray(self.data.astype(dtype), self.indices, self._shape)

    def copyto(self, other):
        if isinstance(other, RowSparseNDArray):
            other.data, other.indices, other._shape = self.data.copy(), self.indices.copy(), self._shape
            return other
        if isinstance(other, NDArray):
            other[:] = self.tostype("default")
            return other
        return self.as_in_context(other)

    def _rebuild(self, data, move=None):
        return RowSparseNDArray(data, self.indices.as_in_context(move) if move is not None else self.indices.copy(), self._shape)

    def _check(self, full):
        idx = self.indices._t.long()
        if idx.numel() != self.data._t.shape[0] or tuple(self.data._t.shape[1:]) != tuple(self._shape[1:]):
            raise MXNetError("row_sparse: data has %s rows of shape %s for %d indices into an array of shape %s"
                             % (self.data._t.shape[0], tuple(self.data._t.shape[1:]), idx.numel(), self._shape))
        if full and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self._shape[0] or bool((idx[1:] <= idx[:-1]).any())):
            raise MXNetError("row_sparse: indices must be strictly increasing and within [0, %d)" % self._shape[0])

    def retain(self, row_ids):
        """sparse_retain: keep only the listed rows (ids need not be present)."""
        ids = row_ids._t.reshape(-1).long() if isinstance(row_ids, NDArray) else torch.as_tensor(row_ids, dtype=torch.int64)
        ids = ids.to(self.indices._t.device)
        keep = torch.isin(self.indices._t.long(), ids)
        return RowSparseNDArray(NDArray(self.data._t[keep]), NDArray(self.indices._t[keep]), self._shape)

    def _set_rows(self, data_t, ids_t):
        self.data, self.indices = NDArray(data_t), NDArray(ids_t)

    # ---- arithmetic that keeps the format (enough for gradient accumulation / scaling)
    def __mul__(self, s): return RowSparseNDArray(self.data * s, self.indices, self._shape)
    __rmul__ = __mul__
    def __truediv__(self, s): return RowSparseNDArray(self.data / s, self.indices, self._shape)
    def __add__(self, other):
        if isinstance(other, RowSparseNDArray):
            return add(self, other)
        return self.tostype("default") + other


class CSRNDArray(BaseSparseNDArray):
    """Compressed sparse rows (``python/mxnet/ndarray/sparse.py`` CSRNDArray :260-560): ``data`` (nnz,), ``indices`` (nnz,) column ids,
    ``indptr`` (rows+1,).  Enough for storage conversion, (de)serialisation, slicing by rows and ``dot(csr, dense)``."""
    stype = "csr"

    def __init__(self, data, indices, indptr, shape):
        self.data, self.indices, self.indptr, self._shape = data, indices, indptr, tuple(int(s) for s in shape)

    shape = property(lambda self: self._shape)
    dtype = property(lambda self: self.data.dtype)
    context = property(lambda self: self.data.context)
    ctx = context

    def __repr__(self):
        return "<CSRNDArray %dx%d @%s nnz=%d>" % (self._shape[0], self._shape[1], self.context, self.data.shape[0])

    def _torch(self):
        """torch CSR view; column indices are sorted per row first (MXNet allows unsorted rows, torch's kernels do not)."""
        ptr, idx, val = self.indptr._t.long(), self.indices._t.long(), self.data._t
        rows = torch.repeat_interleave(torch.arange(self._shape[0], device=ptr.device), ptr[1:] - ptr[:-1])
        order = torch.argsort(rows * self._shape[1] + idx, stable=True)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)      # "Sparse CSR tensor support is in beta state"
            return torch.sparse_csr_tensor(ptr, idx[order], val[order], size=self._shape, check_invariants=False)

    def tostype(self, stype):
        if stype == "csr":
            return self
        dense = NDArray(self._torch().to_dense())
        return dense if stype == "default" else cast_storage(dense, stype)

    todense = lambda self: self.tostype("default")
    def asnumpy(self): return self.tostype("default").asnumpy()
    def copy(self): return CSRNDArray(self.data.copy(), self.indices.copy(), self.indptr.copy(), self._shape)

    def _rebuild(self, data, move=None):
        mv = (lambda x: x.as_in_context(move)) if move is not None else (lambda x: x.copy())
        return CSRNDArray(data, mv(self.indices), mv(self.indptr), self._shape)

    def _check(self, full):
        ptr, idx = self.indptr._t.long(), self.indices._t.long()
        if ptr.numel() != self._shape[0] + 1 or int(ptr[0]) != 0 or int(ptr[-1]) != idx.numel() or idx.numel() != self.data._t.shape[0]:
            raise MXNetError("csr: indptr / indices / data lengths are inconsistent")
        if full and (bool((ptr[1:] < ptr[:-1]).any()) or (idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self._shape[1]))):
            raise MXNetError("csr: indptr must be non-decreasing and column ids within [0, %d)" % self._shape[1])

    def asscipy(self):
        """``scipy.sparse.csr_matrix`` sharing nothing with this array."""
        import scipy.sparse as sp
        return sp.csr_matrix((self.data.asnumpy(), self.indices.asnumpy(), self.indptr.asnumpy()), shape=self._shape)

    def __getitem__(self, key):
        if isinstance(key, int):
            key = slice(key, key + 1)
        start, stop, step = key.indices(self._shape[0])
        if step != 1:
            raise MXNetError("CSRNDArray only supports contiguous row slices")
        ptr = self.indptr._t.long()
        lo, hi = int(ptr[start]), int(ptr[stop])
        return CSRNDArray(NDArray(self.data._t[lo:hi].clone()), NDArray(self.indices._t[lo:hi].clone()), NDArray(ptr[start:stop + 1] - lo),
                          (stop - start, self._shape[1]))


def csr_matrix(arg1, shape=None, ctx=None, dtype=None):
    """``(data, indices, indptr)`` + shape, or a dense array-like."""
    if isinstance(arg1, CSRNDArray):
        return arg1.copy()
    if isinstance(arg1, tuple) and len(arg1) == 3:
        data, indices, indptr = arg1
        mk = lambda x, dt: x if isinstance(x, NDArray) else _dense_array(np.asarray(x), ctx=ctx, dtype=dt)
        if shape is None:
            raise MXNetError("csr_matrix((data, indices, indptr)) needs shape=")
        return CSRNDArray(mk(data, dtype or "float32"), mk(indices, "int64"), mk(indptr, "int64"), shape)
    dense = arg1 if isinstance(arg1, NDArray) else _dense_array(arg1, ctx=ctx, dtype=dtype or "float32")
    return cast_storage(dense, "csr")


def dot(lhs, rhs, transpose_a=False, transpose_b=False):
    """``dot(csr, dense)`` / ``dot(csr.T, dense)`` (src/operator/tensor/dot-inl.h sparse paths); dense operands fall through to ``nd.dot``."""
    if isinstance(lhs, CSRNDArray):
        a = lhs._torch()
        r = rhs._t.t() if transpose_b else rhs._t
        if transpose_a:
            return NDArray(torch.matmul(a.to_dense().t(), r))
        return NDArray(torch.matmul(a, r))
    from . import ndarray as _nd
    return _nd.dot(lhs, rhs, transpose_a, transpose_b)


def row_sparse_array(arg1, shape=None, ctx=None, dtype=None):
    """``(data, indices)`` + shape, a dense array-like (rows that are entirely zero are dropped), or another RowSparseNDArray."""
    if isinstance(arg1, RowSparseNDArray):
        return arg1.copy()
    if isinstance(arg1, tuple) and len(arg1) == 2:
        data, idx = arg1
        data = data if isinstance(data, NDArray) else _dense_array(data, ctx=ctx, dtype=dtype or "float32")
        idx = idx if isinstance(idx, NDArray) else _dense_array(np.asarray(idx, dtype=np.int64), ctx=ctx, dtype="int64")
        order = torch.argsort(idx._t.long())
        data, idx = NDArray(data._t[order]), NDArray(idx._t.long()[order])
        if shape is None:
            raise MXNetError("row_sparse_array((data, indices)) needs shape=")
        return RowSparseNDArray(data, idx, shape)
    dense = arg1 if isinstance(arg1, NDArray) else _dense_array(arg1, ctx=ctx, dtype=dtype or "float32")
    return cast_storage(dense, "row_sparse")


def cast_storage(arr, stype):
    if isinstance(arr, RowSparseNDArray):
        return arr.tostype(stype)
    if isinstance(arr, CSRNDArray):
        return arr.tostype(stype)
    if stype == "default":
        return arr
    if stype == "csr":
        t = arr._t
        if t.dim() != 2:
            raise MXNetError("csr storage needs a 2-D array")
        sp = t.to_sparse_csr()
        return CSRNDArray(NDArray(sp.values().clone()), NDArray(sp.col_indices().clone()), NDArray(sp.crow_indices().clone()), t.shape)
    if stype != "row_sparse":
        raise MXNetError("cast_storage to %s is not supported" % stype)
    t = arr._t
    nz = (t.reshape(t.shape[0], -1) != 0).any(dim=1).nonzero().reshape(-1)
    return RowSparseNDArray(NDArray(_gather(t.contiguous(), nz)), NDArray(nz), t.shape)


def zeros(stype, shape, ctx=None, dtype="float32"):
    from . import ndarray as _nd
    if stype == "default":
        return _nd.zeros(shape, ctx=ctx, dtype=dtype)
    if stype == "csr":
        ctx = ctx or current_context()
        return CSRNDArray(_nd.zeros((0,), ctx=ctx, dtype=dtype), _nd.zeros((0,), ctx=ctx, dtype="int64"), _nd.zeros((shape[0] + 1,), ctx=ctx, dtype="int64"),
                          tuple(shape))
    if stype != "row_sparse":
        raise MXNetError("zeros(%s) is not supported" % stype)
    ctx = ctx or current_context()
    shape = tuple(shape)
    return RowSparseNDArray(_nd.zeros((0,) + shape[1:], ctx=ctx, dtype=dtype), _nd.zeros((0,), ctx=ctx, dtype="int64"), shape)


def add(a: RowSparseNDArray, b: RowSparseNDArray):
    if a.shape != b.shape:
        raise MXNetError("shape mismatch")
    ids = torch.cat([a.indices._t.long(), b.indices._t.long().to(a.indices._t.device)])
    rows = torch.cat([a.data._t, b.data._t.to(a.data._t.device)])
    uniq, inv = torch.unique(ids, sorted=True, return_inverse=True)
    out = torch.zeros((uniq.numel(),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    out.index_add_(0, inv, rows)
    return RowSparseNDArray(NDArray(out), NDArray(uniq), a.shape)


def retain(arr, row_ids):
    return arr.retain(row_ids)


import abc as _abc  # noqa: E402


class BaseSparseNDArray(_abc.ABC):
    """Marker base of the sparse array types (``isinstance(x, BaseSparseNDArray)``)."""


BaseSparseNDArray.register(RowSparseNDArray)
BaseSparseNDArray.register(CSRNDArray)


def array(source_array, ctx=None, dtype=None):
    """Sparse array from a scipy.sparse matrix, another sparse NDArray or a ``(data, indices[, indptr])`` description (sparse.py:1560-1640)."""
    if isinstance(source_array, (RowSparseNDArray, CSRNDArray)):
        return cast_storage(source_array.tostype("default"), source_array.stype)
    try:
        import scipy.sparse as sp
        if sp.issparse(source_array):
            m = source_array.tocsr()
            return csr_matrix((m.data, m.indices, m.indptr), shape=m.shape, ctx=ctx, dtype=dtype)
    except ImportError:
        pass
    raise ValueError("Unexpected source_array type: %s" % type(source_array))


def empty(stype, shape, ctx=None, dtype=None):
    return zeros(stype, shape, ctx=ctx, dtype=dtype or "float32")


def _dense_of(x):
    return x.tostype("default") if isinstance(x, (RowSparseNDArray, CSRNDArray)) else x


def _elemwise(fn, lhs, rhs):
    """Element-wise op; two operands of the same sparse type give that type back, anything else is dense (sparse.py add/subtract/…)."""
    out = fn(_dense_of(lhs), _dense_of(rhs))
    st = getattr(lhs, "stype", "default")
    if st != "default" and st == getattr(rhs, "stype", "default"):
        return cast_storage(out, st)
    return out


def subtract(lhs, rhs): return _elemwise(lambda a, b: a - b, lhs, rhs)
def multiply(lhs, rhs): return _elemwise(lambda a, b: a * b, lhs, rhs)
def divide(lhs, rhs): return _elemwise(lambda a, b: a / b, lhs, rhs)
