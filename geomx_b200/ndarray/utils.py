"""``mx.nd.save`` / ``mx.nd.load`` — byte-compatible with MXNet's NDArray-list format.

Wire format (reference ``src/ndarray/ndarray.cc:1583-1811``, dmlc ``serializer.h``):
  file  := u64 0x112 | u64 0 | u64 n | n*NDArray | u64 m | m*(u64 len | bytes)
  NDArray V2 := u32 0xF993fac9 | i32 stype | [sparse: storage_shape] | shape | ctx | i32 type_flag
                | [aux types/shapes] | raw data | [aux data]
  shape := u32 ndim | i64[ndim];  ctx := i32 dev_type | i32 dev_id
Also accepts V1 (0xF993fac8, no stype) and legacy (magic == ndim, u32 dims) on load.

The hot implementation is native (``csrc/runtime/params_io.cc`` via ``_C``); this module
falls back to the pure-Python codec below when the extension is absent, and the two are
cross-checked in ``tests/test_checkpoint.py``.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from ..base import MXNetError
from ..context import Context, cpu
from .ndarray import NDArray

__all__ = ["save", "load", "load_frombuffer", "save_bytes", "load_bytes", "save_async"]

LIST_MAGIC = 0x112
V2_MAGIC = 0xF993FAC9
V1_MAGIC = 0xF993FAC8
# mshadow type flags (3rdparty/mshadow/mshadow/base.h:302-310)
_FLAG2NP = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_NP2FLAG = {np.dtype(v).name: k for k, v in _FLAG2NP.items()}


def _np_of(a):
    t = a._t.detach()
    if t.dtype == torch.bfloat16:
        t = t.float()
    return np.ascontiguousarray(t.cpu().numpy())


def _enc_shape(shape):
    return struct.pack("<I", len(shape)) + struct.pack("<%dq" % len(shape), *shape)


def _enc_sparse(a) -> bytes:
    """NDArray V2 record of a sparse array (ndarray.cc:1583-1651): stype, storage shape, shape, context, dtype, aux types+shapes, data, aux."""
    from .sparse import CSRNDArray
    csr = isinstance(a, CSRNDArray)
    data = _np_of(a.data)
    auxs = [_np_of(a.indptr).astype(np.int64), _np_of(a.indices).astype(np.int64)] if csr else [_np_of(a.indices).astype(np.int64)]
    ctx = a.context
    out = [struct.pack("<Ii", V2_MAGIC, 2 if csr else 1), _enc_shape(data.shape), _enc_shape(a.shape),
           struct.pack("<ii", ctx.device_typeid, ctx.device_id), struct.pack("<i", _NP2FLAG[data.dtype.name])]
    for x in auxs:
        out.append(struct.pack("<i", 6) + _enc_shape(x.shape))
    out.append(data.tobytes())
    out += [x.tobytes() for x in auxs]
    return b"".join(out)


def _enc_array(a) -> bytes:
    if getattr(a, "stype", "default") in ("row_sparse", "csr"):
        return _enc_sparse(a)
    t = a._t.detach()
    if t.dtype == torch.bfloat16:
        t = t.float()
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    ctx = a.context
    npa = np.ascontiguousarray(t.cpu().numpy())
    if npa.dtype.name not in _NP2FLAG:
        raise MXNetError("dtype %s cannot be serialised" % npa.dtype)
    out = [struct.pack("<Ii", V2_MAGIC, 0)]
    out.append(struct.pack("<I", npa.ndim) + struct.pack("<%dq" % npa.ndim, *npa.shape))
    if npa.ndim == 0 or npa.size == 0:
        # MXNet writes nothing after an empty shape ("none" array)
        if npa.ndim == 0:
            return b"".join(out)
    out.append(struct.pack("<ii", ctx.device_typeid, ctx.device_id))
    out.append(struct.pack("<i", _NP2FLAG[npa.dtype.name]))
    out.append(npa.tobytes())
    return b"".join(out)


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf); self.o = 0

    def take(self, fmt):
        n = struct.calcsize(fmt)
        if self.o + n > len(self.b):
            raise MXNetError("truncated NDArray file")
        v = struct.unpack_from(fmt, self.b, self.o); self.o += n
        return v

    def raw(self, n):
        if self.o + n > len(self.b):
            raise MXNetError("truncated NDArray file")
        v = self.b[self.o:self.o + n]; self.o += n
        return v


def _dec_shape64(r):
    (ndim,) = r.take("<I")
    return tuple(r.take("<%dq" % ndim)) if ndim else ()


def _dec_array(r: _Reader, restore_ctx=False) -> NDArray:
    (magic,) = r.take("<I")
    if magic == V2_MAGIC:
        (stype,) = r.take("<i")
        nad = {0: 0, 1: 1, 2: 2}.get(stype)
        if nad is None:
            raise MXNetError("unknown storage type %d" % stype)
        sshape = _dec_shape64(r) if nad else None
        shape = _dec_shape64(r)
        if len(shape) == 0:
            return NDArray(torch.zeros(0))
        dev_type, dev_id = r.take("<ii")
        (flag,) = r.take("<i")
        aux = []
        for _ in range(nad):
            (aflag,) = r.take("<i"); ashape = _dec_shape64(r); aux.append((aflag, ashape))
        dshape = sshape if nad else shape
        dt = np.dtype(_FLAG2NP[flag])
        n = int(np.prod(dshape)) if len(dshape) else 1
        data = np.frombuffer(r.raw(n * dt.itemsize), dtype=dt).reshape(dshape).copy()
        auxd = []
        for aflag, ashape in aux:
            adt = np.dtype(_FLAG2NP[aflag]); an = int(np.prod(ashape)) if len(ashape) else 0
            auxd.append(np.frombuffer(r.raw(an * adt.itemsize), dtype=adt).reshape(ashape).copy())
        if stype == 1:  # row_sparse: aux = (row ids,)
            from .sparse import RowSparseNDArray
            return RowSparseNDArray(NDArray(torch.from_numpy(data)), NDArray(torch.from_numpy(auxd[0].astype(np.int64))), shape)
        if stype == 2:  # csr: aux = (indptr, indices)
            from .sparse import CSRNDArray
            return CSRNDArray(NDArray(torch.from_numpy(data)), NDArray(torch.from_numpy(auxd[1].astype(np.int64))),
                              NDArray(torch.from_numpy(auxd[0].astype(np.int64))), shape)
    else:
        if magic == V1_MAGIC:
            shape = _dec_shape64(r)
        else:  # legacy: magic is ndim, dims are u32
            ndim = magic
            shape = tuple(r.take("<%dI" % ndim)) if ndim else ()
        if len(shape) == 0:
            return NDArray(torch.zeros(0))
        dev_type, dev_id = r.take("<ii")
        (flag,) = r.take("<i")
        dt = np.dtype(_FLAG2NP[flag]); n = int(np.prod(shape))
        data = np.frombuffer(r.raw(n * dt.itemsize), dtype=dt).reshape(shape).copy()
    t = torch.from_numpy(data)
    ctx = cpu()
    if restore_ctx and dev_type == 2 and torch.cuda.is_available() and dev_id < torch.cuda.device_count():
        ctx = Context("gpu", dev_id); t = t.to(ctx.torch_device)
    return NDArray(t, ctx)


def save_bytes(data) -> bytes:
    if isinstance(data, NDArray) or getattr(data, "stype", None) in ("row_sparse", "csr"):
        data = [data]
    if isinstance(data, dict):
        names, arrays = list(data.keys()), list(data.values())
    else:
        names, arrays = [], list(data)
    for a in arrays:
        if not isinstance(a, NDArray) and getattr(a, "stype", None) not in ("row_sparse", "csr"):
            raise MXNetError("save only accepts NDArray, list of NDArray or dict of str->NDArray")
    out = [struct.pack("<QQQ", LIST_MAGIC, 0, len(arrays))]
    out += [_enc_array(a) for a in arrays]
    out.append(struct.pack("<Q", len(names)))
    for n in names:
        b = n.encode("utf-8"); out.append(struct.pack("<Q", len(b)) + b)
    return b"".join(out)


def load_bytes(buf, restore_ctx=True):
    r = _Reader(buf)
    magic, _res = r.take("<QQ")
    if magic != LIST_MAGIC:
        raise MXNetError("Invalid NDArray file format")
    (n,) = r.take("<Q")
    arrays = [_dec_array(r, restore_ctx) for _ in range(n)]
    (m,) = r.take("<Q")
    names = []
    for _ in range(m):
        (ln,) = r.take("<Q"); names.append(bytes(r.raw(ln)).decode("utf-8"))
    if m and m != n:
        raise MXNetError("Invalid NDArray file format")
    return dict(zip(names, arrays)) if m else arrays


def save(fname, data):
    """Save NDArray / list / dict to ``fname`` in MXNet ``.params`` format."""
    with open(fname, "wb") as f:
        f.write(save_bytes(data))


_path_vars = {}


def save_async(fname, data):
    """``save`` without blocking the training loop: the arrays are snapshotted now (device -> host copy, so later in-place updates do not
    leak into the file), serialisation + file IO run on the host dependency engine (``mx.engine``).  Writes to the same path are ordered by
    the path's engine variable; ``mx.nd.waitall()`` (or ``mx.engine.wait_all()``) waits for them.  Falls back to a synchronous ``save``
    when the native runtime is unavailable."""
    from .. import engine, runtime
    if not runtime.available():
        return save(fname, data)
    if isinstance(data, NDArray):
        snap = NDArray(data._t.detach().to("cpu", copy=True))
    elif isinstance(data, dict):
        snap = {k: NDArray(v._t.detach().to("cpu", copy=True)) for k, v in data.items()}
    else:
        snap = [NDArray(v._t.detach().to("cpu", copy=True)) for v in data]
    path = str(fname)
    var = _path_vars.get(path)
    if var is None:
        var = _path_vars[path] = engine.new_variable()

    def write():
        tmp = path + ".tmp%d" % id(snap)
        with open(tmp, "wb") as f:
            f.write(save_bytes(snap))
        import os
        os.replace(tmp, path)                  # readers never observe a half-written checkpoint
    engine.push(write, mutable_vars=[var], name="nd.save_async")


def load(fname):
    with open(fname, "rb") as f:
        return load_bytes(f.read())


def load_frombuffer(buf):
    """``mx.nd.load`` from an in-memory ``.params`` image (ndarray/utils.py:185-220)."""
    return load_bytes(buf)
