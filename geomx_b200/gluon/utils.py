"""``gluon.utils``: split_data / split_and_load / clip_global_norm / check_sha1.
Parity: ``python/mxnet/gluon/utils.py:91`` (split_and_load)."""
from __future__ import annotations

import hashlib
import math

import torch

from ..context import Context
from ..ndarray import NDArray, array

__all__ = ["split_data", "split_and_load", "clip_global_norm", "check_sha1", "download"]


def split_data(data, num_slice, batch_axis=0, even_split=True):
    size = data.shape[batch_axis]
    if even_split and size % num_slice != 0:
        raise ValueError("data with shape %s cannot be evenly split into %d slices along axis %d. "
                         "Use a batch size that's multiple of %d or set even_split=False." % (str(data.shape), num_slice, batch_axis, num_slice))
    step = size // num_slice
    if not even_split and size < num_slice:
        step, num_slice = 1, size
    t = data._t
    out = []
    for i in range(num_slice):
        lo = i * step; hi = (i + 1) * step if i < num_slice - 1 else size
        out.append(NDArray(t.narrow(batch_axis, lo, hi - lo)))
    return out


def split_and_load(data, ctx_list, batch_axis=0, even_split=True):
    if not isinstance(data, NDArray):
        data = array(data, ctx=ctx_list[0])
    if isinstance(ctx_list, Context):
        ctx_list = [ctx_list]
    if len(ctx_list) == 1:
        return [data.as_in_context(ctx_list[0])]
    slices = split_data(data, len(ctx_list), batch_axis, even_split)
    return [s.as_in_context(c) for s, c in zip(slices, ctx_list)]


def clip_global_norm(arrays, max_norm, check_isfinite=True):
    assert len(arrays) > 0
    total = torch.sqrt(sum((a._t.float().square().sum() for a in arrays)))
    total_f = float(total)
    if check_isfinite and not math.isfinite(total_f):
        import warnings
        warnings.warn(UserWarning("nan or inf is detected. Clipping results will be undefined."), stacklevel=2)
    scale = max_norm / (total_f + 1e-8)
    if scale < 1.0:
        for a in arrays:
            a._t.mul_(scale)
    return total_f


def check_sha1(filename, sha1_hash):
    sha1 = hashlib.sha1()
    with open(filename, "rb") as f:
        while True:
            data = f.read(1048576)
            if not data:
                break
            sha1.update(data)
    return sha1.hexdigest() == sha1_hash


def download(url, path=None, overwrite=False, sha1_hash=None, **kw):
    raise RuntimeError("no network egress in this environment; place files under the dataset root manually "
                       "or use synthetic=True datasets")


from .block import _HookHandle as HookHandle  # noqa: E402,F401  (gluon/utils.py HookHandle: returned by Block.register_forward_*hook)
