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


class RowSparseNDArray:
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
    if stype == "default":
        return arr
    if stype != "row_sparse":
        raise MXNetError("cast_storage to %s is not supported" % stype)
    t = arr._t
    nz = (t.reshape(t.shape[0], -1) != 0).any(dim=1).nonzero().reshape(-1)
    return RowSparseNDArray(NDArray(_gather(t.contiguous(), nz)), NDArray(nz), t.shape)


def zeros(stype, shape, ctx=None, dtype="float32"):
    from . import ndarray as _nd
    if stype == "default":
        return _nd.zeros(shape, ctx=ctx, dtype=dtype)
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
