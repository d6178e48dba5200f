"""The wider ``mx.nd.*`` operator library: the tensor / nn / linalg / sequence / random-free operator families of the reference exposed as
imperative functions over NDArray.

Parity (by family, ``src/operator/``): ``tensor/elemwise_unary_op_{basic,trig}`` + ``elemwise_binary_{op,scalar_op,broadcast_op}*``
(math, comparison, logic, broadcast_*), ``tensor/broadcast_reduce_op_{value,index}`` (prod, nansum, argmax_channel, norm…),
``tensor/matrix_op`` (slice, slice_axis, slice_like, expand_dims, squeeze, swapaxes, flip/reverse, repeat, tile, split, space/depth),
``tensor/indexing_op`` (take, batch_take, gather_nd, scatter_nd, Embedding, one_hot), ``tensor/init_op`` (eye, linspace, full_like),
``tensor/ordering_op`` (topk, sort, argsort), ``tensor/dot`` (batch_dot), ``tensor/la_op`` (linalg gemm/gemm2/potrf/trsm/syrk/sumlogdiag…),
``tensor/control_flow_op`` (where), ``sequence_{mask,last,reverse}``, ``nn/`` functional forms (FullyConnected, Convolution, Deconvolution,
Pooling, Activation, LeakyReLU, BatchNorm, LayerNorm, InstanceNorm, L2Normalization, LRN, Dropout, SoftmaxActivation, UpSampling,
Embedding, Pad), ``loss_binary_op`` (softmax_cross_entropy), ``regression_output``/``make_loss`` helpers, ``contrib`` (div_sqrt_dim,
AdaptiveAvgPooling2D, BilinearResize2D, quadratic, index_copy, arange_like).

All functions are differentiable through the autograd tape (they are PyTorch expressions on the backing tensors); Conv/FC/Pool/BN route
through ``ops.functional`` and therefore through the native sm_100a kernels on CUDA.  Names and argument conventions follow MXNet."""
from __future__ import annotations

import builtins as _bi
import math as _math

import numpy as _np
import torch
import torch.nn.functional as TF

from ..ops import functional as OF
from .ndarray import NDArray, _ctx_of, torch_dtype

_W = NDArray


def _t(x):
    return x._t if isinstance(x, NDArray) else x


def _like(x, ref):
    return x._t if isinstance(x, NDArray) else torch.as_tensor(x, dtype=ref.dtype, device=ref.device)


def _axes(axis):
    return None if axis is None or axis == () else (axis if isinstance(axis, int) else tuple(axis))


# ------------------------------------------------------------------------------------------------ unary math
def _unary(fn):
    return lambda data, **kw: _W(fn(_t(data)))


negative = _unary(torch.neg); reciprocal = _unary(torch.reciprocal); rsqrt = _unary(torch.rsqrt); cbrt = _unary(lambda t: torch.sign(t) * t.abs().pow(1 / 3))
rcbrt = _unary(lambda t: 1.0 / (torch.sign(t) * t.abs().pow(1 / 3))); log2 = _unary(torch.log2); log10 = _unary(torch.log10); log1p = _unary(torch.log1p)
expm1 = _unary(torch.expm1); sin = _unary(torch.sin); cos = _unary(torch.cos); tan = _unary(torch.tan); arcsin = _unary(torch.asin)
arccos = _unary(torch.acos); arctan = _unary(torch.atan); sinh = _unary(torch.sinh); cosh = _unary(torch.cosh); arcsinh = _unary(torch.asinh)
arccosh = _unary(torch.acosh); arctanh = _unary(torch.atanh); degrees = _unary(torch.rad2deg); radians = _unary(torch.deg2rad)
floor = _unary(torch.floor); ceil = _unary(torch.ceil); trunc = _unary(torch.trunc); rint = _unary(torch.round); round = _unary(lambda t: torch.sign(t) * torch.floor(t.abs() + 0.5))
fix = _unary(torch.trunc); gamma = _unary(lambda t: torch.exp(torch.lgamma(t))); gammaln = _unary(torch.lgamma); erf = _unary(torch.erf)
erfinv = _unary(torch.erfinv); softsign = _unary(TF.softsign); logical_not = _unary(lambda t: (t == 0).to(t.dtype))
identity = _unary(lambda t: t.clone()); stop_gradient = BlockGrad = _unary(lambda t: t.detach()); make_loss = MakeLoss = _unary(lambda t: t)
ones_like_op = _unary(torch.ones_like)


def hard_sigmoid(data, alpha=0.2, beta=0.5): return _W(torch.clamp(_t(data) * alpha + beta, 0.0, 1.0))
def smooth_l1(data, scalar=1.0):
    t, s2 = _t(data), scalar * scalar
    return _W(torch.where(t.abs() < 1.0 / s2, 0.5 * s2 * t * t, t.abs() - 0.5 / s2))


# ------------------------------------------------------------------------------------------------ binary (elementwise + broadcast + scalar)
def _binary(fn, cast_bool=False):
    def op(lhs, rhs, **kw):
        a = _t(lhs) if isinstance(lhs, NDArray) else None
        b = _t(rhs) if isinstance(rhs, NDArray) else None
        ref = a if a is not None else b
        a = a if a is not None else torch.as_tensor(lhs, dtype=ref.dtype, device=ref.device)
        b = b if b is not None else torch.as_tensor(rhs, dtype=ref.dtype, device=ref.device)
        out = fn(a, b)
        return _W(out.to(ref.dtype) if cast_bool else out)
    return op


add = broadcast_add = broadcast_plus = elemwise_add_op = _binary(torch.add); subtract = broadcast_sub = broadcast_minus = elemwise_sub = _binary(torch.sub)
multiply = broadcast_mul = elemwise_mul = _binary(torch.mul); divide = broadcast_div = elemwise_div = _binary(torch.div)
modulo = broadcast_mod = _binary(torch.fmod); power = broadcast_power = _binary(torch.pow); hypot = broadcast_hypot = _binary(torch.hypot)
broadcast_maximum = _binary(torch.maximum); broadcast_minimum = _binary(torch.minimum)
equal = broadcast_equal = _binary(torch.eq, True); not_equal = broadcast_not_equal = _binary(torch.ne, True)
greater = broadcast_greater = _binary(torch.gt, True); greater_equal = broadcast_greater_equal = _binary(torch.ge, True)
lesser = broadcast_lesser = _binary(torch.lt, True); lesser_equal = broadcast_lesser_equal = _binary(torch.le, True)
logical_and = broadcast_logical_and = _binary(lambda a, b: (a != 0) & (b != 0), True)
logical_or = broadcast_logical_or = _binary(lambda a, b: (a != 0) | (b != 0), True)
logical_xor = broadcast_logical_xor = _binary(lambda a, b: (a != 0) ^ (b != 0), True)


def broadcast_like(lhs, rhs): return _W(_t(lhs).expand_as(_t(rhs)))
def broadcast_axis(data, axis=0, size=1):
    t = _t(data); shape = list(t.shape)
    for a, s in zip([axis] if isinstance(axis, int) else axis, [size] if isinstance(size, int) else size):
        shape[a] = s
    return _W(t.expand(*shape))
broadcast_axes = broadcast_axis


# ------------------------------------------------------------------------------------------------ reductions
def prod(data, axis=None, keepdims=False):
    t = _t(data); ax = _axes(axis)
    if ax is None:
        return _W(t.prod().reshape(1) if not keepdims else t.prod().reshape([1] * t.dim()))
    for a in sorted([ax] if isinstance(ax, int) else ax, reverse=True):
        t = t.prod(dim=a, keepdim=keepdims)
    return _W(t)


def nansum(data, axis=None, keepdims=False): return _W(torch.nansum(_t(data), dim=_axes(axis), keepdim=keepdims) if axis is not None else torch.nansum(_t(data)).reshape(1))
def nanprod(data, axis=None, keepdims=False): return prod(_W(torch.nan_to_num(_t(data), nan=1.0)), axis, keepdims)
def argmax_channel(data): return _W(_t(data).argmax(dim=1).to(_t(data).dtype))
def square_sum(data, axis=None, keepdims=False): return _W((_t(data) ** 2).sum(dim=_axes(axis), keepdim=keepdims) if axis is not None else (_t(data) ** 2).sum().reshape(1))
def L2Normalization(data, eps=1e-10, mode="instance"):
    t = _t(data)
    dims = {"instance": tuple(range(1, t.dim())), "channel": (1,), "spatial": tuple(range(2, t.dim()))}[mode]
    return _W(t / torch.sqrt((t * t).sum(dim=dims, keepdim=True) + eps))


# ------------------------------------------------------------------------------------------------ shape manipulation
def expand_dims(data, axis): return _W(_t(data).unsqueeze(axis))
def squeeze(data, axis=None): return _W(_t(data).squeeze() if axis is None else _t(data).squeeze(axis))
def swapaxes(data, dim1=0, dim2=0): return _W(_t(data).transpose(dim1, dim2))
SwapAxis = swapaxes
def flip(data, axis): return _W(_t(data).flip([axis] if isinstance(axis, int) else list(axis)))
reverse = flip
def repeat(data, repeats, axis=None): return _W(_t(data).repeat_interleave(repeats) if axis is None else _t(data).repeat_interleave(repeats, dim=axis))
def Flatten(data): return _W(_t(data).flatten(1))
def Reshape(data, shape, reverse=False): return data.reshape(shape)
def reshape_like(lhs, rhs): return _W(_t(lhs).reshape(_t(rhs).shape))
def shape_array(data): return _W(torch.tensor(list(_t(data).shape), dtype=torch.int64, device=_t(data).device))
def size_array(data): return _W(torch.tensor([_t(data).numel()], dtype=torch.int64, device=_t(data).device))
def Cast(data, dtype): return data.astype(dtype)


def slice(data, begin, end, step=None):
    t = _t(data)
    idx = []
    for i in range(len(begin)):
        st = None if step is None or step[i] is None else step[i]
        idx.append(_bi.slice(begin[i], end[i], st))
    return _W(t[tuple(idx)])


def slice_axis(data, axis, begin, end):
    t = _t(data)
    end = t.shape[axis] if end is None else end
    return _W(t.narrow(axis, begin if begin >= 0 else t.shape[axis] + begin, (end if end >= 0 else t.shape[axis] + end) - (begin if begin >= 0 else t.shape[axis] + begin)))


def slice_like(data, shape_like, axes=()):
    t, r = _t(data), _t(shape_like)
    idx = [_bi.slice(None)] * t.dim()
    for a in (axes or range(min(t.dim(), r.dim()))):
        idx[a] = _bi.slice(0, r.shape[a])
    return _W(t[tuple(idx)])


def split(data, num_outputs, axis=1, squeeze_axis=False):
    parts = torch.chunk(_t(data), num_outputs, dim=axis)
    outs = [_W(p.squeeze(axis) if squeeze_axis else p) for p in parts]
    return outs[0] if num_outputs == 1 else outs
SliceChannel = split


def Concat(*data, dim=1, num_args=None): return _W(torch.cat([_t(d) for d in data], dim=dim))
def depth_to_space(data, block_size): return _W(TF.pixel_shuffle(_t(data), block_size))
def space_to_depth(data, block_size): return _W(TF.pixel_unshuffle(_t(data), block_size))
def diag(data, k=0): return _W(torch.diag(_t(data), k) if _t(data).dim() <= 2 else torch.diagonal(_t(data), k, -2, -1))


def Pad(data, mode="constant", pad_width=(), constant_value=0.0):
    pw = list(pad_width)
    pairs = [(pw[i], pw[i + 1]) for i in range(0, len(pw), 2)]
    flat = []
    for lo, hi in reversed(pairs):
        flat += [lo, hi]
    return _W(TF.pad(_t(data), flat, mode={"constant": "constant", "edge": "replicate", "reflect": "reflect"}[mode], value=constant_value if mode == "constant" else 0.0))
pad = Pad


# ------------------------------------------------------------------------------------------------ creation
def eye(N, M=0, k=0, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    return _W(torch.diag(torch.ones(max(0, min(N, (M or N) - k) if k >= 0 else min(N + k, M or N)), dtype=torch_dtype(dtype)), k)[:N, :(M or N)].contiguous().to(ctx.torch_device)
              if k else torch.eye(N, M or N, dtype=torch_dtype(dtype), device=ctx.torch_device), ctx)


def linspace(start, stop, num, endpoint=True, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    t = torch.linspace(start, stop, num if endpoint else num + 1, dtype=torch_dtype(dtype), device=ctx.torch_device)
    return _W(t if endpoint else t[:-1], ctx)


def full_like(data, fill_value): return _W(torch.full_like(_t(data), fill_value))
def arange_like(data, start=0.0, step=1.0, axis=None):
    t = _t(data)
    n = t.numel() if axis is None else t.shape[axis]
    r = torch.arange(n, dtype=t.dtype, device=t.device) * step + start
    return _W(r.reshape(t.shape) if axis is None else r)


# ------------------------------------------------------------------------------------------------ indexing
def take(a, indices, axis=0, mode="clip"):
    t, idx = _t(a), _t(indices).long()
    n = t.shape[axis]
    idx = idx.clamp(0, n - 1) if mode == "clip" else idx % n
    return _W(torch.index_select(t, axis, idx.reshape(-1)).reshape(t.shape[:axis] + tuple(idx.shape) + t.shape[axis + 1:]))


def batch_take(a, indices): return _W(_t(a).gather(1, _t(indices).long().view(-1, 1)).view(-1))
def gather_nd(data, indices):
    t, idx = _t(data), _t(indices).long()
    return _W(t[tuple(idx[i] for i in range(idx.shape[0]))])


def scatter_nd(data, indices, shape):
    t, idx = _t(data), _t(indices).long()
    out = torch.zeros(tuple(shape), dtype=t.dtype, device=t.device)
    out[tuple(idx[i] for i in range(idx.shape[0]))] = t
    return _W(out)


def Embedding(data, weight, input_dim=None, output_dim=None, sparse_grad=False): return _W(TF.embedding(_t(data).long(), _t(weight)))
def index_copy(old, index, new):
    out = _t(old).clone(); out.index_copy_(0, _t(index).long(), _t(new))
    return _W(out)


def SequenceMask(data, sequence_length=None, use_sequence_length=False, value=0.0, axis=0):
    t = _t(data)
    if not use_sequence_length or sequence_length is None:
        return _W(t.clone())
    T = t.shape[axis]
    steps = torch.arange(T, device=t.device).view([-1 if i == axis else 1 for i in range(t.dim())])
    lens = _t(sequence_length).view([-1 if i == (1 - axis) else 1 for i in range(t.dim())])
    return _W(torch.where(steps < lens, t, torch.full_like(t, value)))


def SequenceLast(data, sequence_length=None, use_sequence_length=False, axis=0):
    t = _t(data)
    if not use_sequence_length or sequence_length is None:
        return _W(t.select(axis, t.shape[axis] - 1))
    idx = (_t(sequence_length).long() - 1).clamp_min(0)
    tt = t if axis == 0 else t.transpose(0, 1)
    return _W(tt[idx, torch.arange(tt.shape[1], device=t.device)])


def SequenceReverse(data, sequence_length=None, use_sequence_length=False, axis=0):
    t = _t(data)
    if not use_sequence_length or sequence_length is None:
        return _W(t.flip(0))
    T = t.shape[0]
    lens = _t(sequence_length).long()
    steps = torch.arange(T, device=t.device).view(-1, 1)
    src = torch.where(steps < lens.view(1, -1), lens.view(1, -1) - 1 - steps, steps)
    return _W(t[src, torch.arange(t.shape[1], device=t.device).view(1, -1)])


# ------------------------------------------------------------------------------------------------ linear algebra
def batch_dot(lhs, rhs, transpose_a=False, transpose_b=False):
    a, b = _t(lhs), _t(rhs)
    return _W(torch.bmm(a.transpose(1, 2) if transpose_a else a, b.transpose(1, 2) if transpose_b else b))


def linalg_gemm2(A, B, transpose_a=False, transpose_b=False, alpha=1.0):
    a, b = _t(A), _t(B)
    return _W(alpha * torch.matmul(a.transpose(-1, -2) if transpose_a else a, b.transpose(-1, -2) if transpose_b else b))


def linalg_gemm(A, B, C, transpose_a=False, transpose_b=False, alpha=1.0, beta=1.0): return _W(_t(linalg_gemm2(A, B, transpose_a, transpose_b, alpha)) + beta * _t(C))
def linalg_potrf(A): return _W(torch.linalg.cholesky(_t(A)))
def linalg_potri(A):
    L = _t(A)
    return _W(torch.cholesky_inverse(L))
def linalg_trsm(A, B, transpose=False, rightside=False, lower=True, alpha=1.0):
    a, b = _t(A), _t(B)
    a = a.transpose(-1, -2) if transpose else a
    low = lower != transpose
    if rightside:
        return _W(alpha * torch.linalg.solve_triangular(a, b, upper=not low, left=False))
    return _W(alpha * torch.linalg.solve_triangular(a, b, upper=not low))
def linalg_trmm(A, B, transpose=False, rightside=False, lower=True, alpha=1.0):
    a = torch.tril(_t(A)) if lower else torch.triu(_t(A))
    a = a.transpose(-1, -2) if transpose else a
    return _W(alpha * (torch.matmul(_t(B), a) if rightside else torch.matmul(a, _t(B))))
def linalg_syrk(A, transpose=False, alpha=1.0):
    a = _t(A)
    return _W(alpha * (torch.matmul(a.transpose(-1, -2), a) if transpose else torch.matmul(a, a.transpose(-1, -2))))
def linalg_sumlogdiag(A): return _W(torch.log(torch.diagonal(_t(A), dim1=-2, dim2=-1)).sum(-1))
def linalg_inverse(A): return _W(torch.linalg.inv(_t(A)))
def linalg_det(A): return _W(torch.linalg.det(_t(A)))
def linalg_extractdiag(A, offset=0): return _W(torch.diagonal(_t(A), offset, -2, -1))
def linalg_makediag(A, offset=0): return _W(torch.diag_embed(_t(A), offset))
def khatri_rao(*mats):
    out = _t(mats[0])
    for m in mats[1:]:
        out = torch.einsum("ik,jk->ijk", out, _t(m)).reshape(-1, out.shape[1])
    return _W(out)


# ------------------------------------------------------------------------------------------------ nn functional forms
def FullyConnected(data, weight, bias=None, num_hidden=None, no_bias=False, flatten=True):
    return _W(OF.dense(_t(data), _t(weight), None if no_bias or bias is None else _t(bias), None, flatten))


def Convolution(data, weight, bias=None, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0), num_filter=None, num_group=1, no_bias=False, layout=None):
    t = _t(data)
    if t.dim() == 4:
        return _W(OF.conv2d(t, _t(weight), None if no_bias or bias is None else _t(bias), tuple(stride), tuple(pad), tuple(dilate), num_group))
    fn = {3: TF.conv1d, 5: TF.conv3d}[t.dim()]
    return _W(fn(t, _t(weight), None if no_bias or bias is None else _t(bias), tuple(stride), tuple(pad), tuple(dilate), num_group))


def Deconvolution(data, weight, bias=None, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0), adj=(0, 0), num_filter=None, num_group=1, no_bias=True):
    fn = {3: TF.conv_transpose1d, 4: TF.conv_transpose2d, 5: TF.conv_transpose3d}[_t(data).dim()]
    return _W(fn(_t(data), _t(weight), None if no_bias or bias is None else _t(bias), tuple(stride), tuple(pad), tuple(adj), num_group, tuple(dilate)))


def Pooling(data, kernel=(2, 2), pool_type="max", stride=None, pad=(0, 0), global_pool=False, pooling_convention="valid", count_include_pad=True):
    t = _t(data)
    k = tuple(t.shape[2:]) if global_pool else tuple(kernel)
    st = tuple(stride) if stride else k
    ceil_mode = pooling_convention == "full"
    if t.dim() == 4 and pool_type in ("max", "avg"):
        return _W(OF.max_pool2d(t, k, st, tuple(pad), ceil_mode) if pool_type == "max" else OF.avg_pool2d(t, k, st, tuple(pad), ceil_mode, count_include_pad))
    if pool_type == "sum":
        return _W(TF.avg_pool2d(t, k, st, tuple(pad), ceil_mode, True) * float(_np.prod(k)))
    if pool_type == "lp":
        return _W(TF.lp_pool2d(t, 2.0, k, st, ceil_mode))
    fn = {("max", 3): TF.max_pool1d, ("avg", 3): TF.avg_pool1d, ("max", 5): TF.max_pool3d, ("avg", 5): TF.avg_pool3d}[(pool_type, t.dim())]
    return _W(fn(t, k, st, tuple(pad), ceil_mode=ceil_mode))


def Activation(data, act_type="relu"): return _W(OF.activation(_t(data), act_type))
def LeakyReLU(data, gamma=None, act_type="leaky", slope=0.25, lower_bound=0.125, upper_bound=0.334):
    t = _t(data)
    if act_type == "leaky": return _W(TF.leaky_relu(t, slope))
    if act_type == "elu": return _W(TF.elu(t, slope))
    if act_type == "selu": return _W(TF.selu(t))
    if act_type == "gelu": return _W(TF.gelu(t))
    if act_type == "prelu": return _W(TF.prelu(t, _t(gamma)))
    if act_type == "rrelu": return _W(TF.rrelu(t, lower_bound, upper_bound, training=False))
    raise ValueError(act_type)


def BatchNorm(data, gamma, beta, moving_mean, moving_var, eps=1e-3, momentum=0.9, fix_gamma=True, use_global_stats=False, axis=1, training=None):
    from .. import autograd
    tr = (autograd.is_training() if training is None else training) and not use_global_stats
    g = torch.ones_like(_t(gamma)) if fix_gamma else _t(gamma)
    return _W(OF.batch_norm(_t(data), g, _t(beta), _t(moving_mean), _t(moving_var), tr, momentum, eps, axis))


def LayerNorm(data, gamma, beta, axis=-1, eps=1e-5): return _W(OF.layer_norm(_t(data), _t(gamma), _t(beta), axis, eps))
def InstanceNorm(data, gamma, beta, eps=1e-3): return _W(TF.instance_norm(_t(data), weight=_t(gamma), bias=_t(beta), eps=eps))
def LRN(data, alpha=1e-4, beta=0.75, knorm=2.0, nsize=5): return _W(TF.local_response_norm(_t(data), nsize, alpha * nsize / nsize, beta, knorm))
def Dropout(data, p=0.5, mode="training"):
    from .. import autograd
    return _W(OF.dropout(_t(data), p, mode == "always" or autograd.is_training()))
def SoftmaxActivation(data, mode="instance"): return _W(torch.softmax(_t(data), dim=1 if mode == "channel" else -1))
def softmin(data, axis=-1): return _W(torch.softmax(-_t(data), dim=axis))
def softmax_cross_entropy(data, label): return _W(TF.cross_entropy(_t(data), _t(label).long(), reduction="sum").reshape(1))
def UpSampling(data, scale=2, sample_type="nearest", num_args=1): return _W(TF.interpolate(_t(data), scale_factor=scale, mode="nearest" if sample_type == "nearest" else "bilinear"))
def BilinearResize2D(data, height, width): return _W(TF.interpolate(_t(data), size=(height, width), mode="bilinear", align_corners=True))
def AdaptiveAvgPooling2D(data, output_size=1): return _W(TF.adaptive_avg_pool2d(_t(data), output_size))
def div_sqrt_dim(data): return _W(_t(data) / _math.sqrt(_t(data).shape[-1]))
def quadratic(data, a=0.0, b=0.0, c=0.0): return _W(a * _t(data) ** 2 + b * _t(data) + c)
def LinearRegressionOutput(data, label=None, grad_scale=1.0): return _W(_t(data).clone())
def LogisticRegressionOutput(data, label=None, grad_scale=1.0): return _W(torch.sigmoid(_t(data)))
def MAERegressionOutput(data, label=None, grad_scale=1.0): return _W(_t(data).clone())
def SoftmaxOutput(data, label=None, **kw): return _W(torch.softmax(_t(data), dim=1))
def ctc_loss(data, label, data_lengths=None, label_lengths=None, use_data_lengths=False, use_label_lengths=False, blank_label="first"):
    t, lab = _t(data), _t(label).long()
    T, N = t.shape[0], t.shape[1]
    dl = _t(data_lengths).long() if use_data_lengths else torch.full((N,), T, dtype=torch.long)
    ll = _t(label_lengths).long() if use_label_lengths else (lab != (0 if blank_label == "first" else -1)).sum(1)
    blank = 0 if blank_label == "first" else t.shape[2] - 1
    return _W(TF.ctc_loss(torch.log_softmax(t, 2), lab, dl, ll, blank=blank, reduction="none"))
CTCLoss = ctc_loss



# ------------------------------------------------------------------------------------------------ spatial transformer family
def GridGenerator(data, transform_type="affine", target_shape=(0, 0)):
    """Sampling grid in [-1, 1] (``grid_generator-inl.h``): ``affine`` — ``data [B, 6]`` → ``[B, 2, H, W]`` (x, y); ``warp`` — ``data [B, 2, H, W]``
    optical flow in pixels added to the identity grid."""
    x = _t(data)
    if transform_type == "affine":
        H, W = int(target_shape[0]), int(target_shape[1])
        g = TF.affine_grid(x.reshape(-1, 2, 3), (x.shape[0], 1, H, W), align_corners=True)     # [B, H, W, 2]
        return _W(g.permute(0, 3, 1, 2))
    B, _, H, W = x.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=x.dtype, device=x.device), torch.arange(W, dtype=x.dtype, device=x.device), indexing="ij")
    gx = (xs[None] + x[:, 0]) / _bi.max(W - 1, 1) * 2 - 1
    gy = (ys[None] + x[:, 1]) / _bi.max(H - 1, 1) * 2 - 1
    return _W(torch.stack([gx, gy], 1))


def BilinearSampler(data, grid):
    """Bilinear sampling of ``data [B, C, H, W]`` at ``grid [B, 2, Ho, Wo]`` (x, y in [-1, 1]); outside → 0 (``bilinear_sampler-inl.h``)."""
    return _W(TF.grid_sample(_t(data), _t(grid).permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros", align_corners=True))


def SpatialTransformer(data, loc, target_shape=(0, 0), transform_type="affine", sampler_type="bilinear"):
    """Affine spatial transformer network op = GridGenerator + BilinearSampler (``spatial_transformer-inl.h``)."""
    return BilinearSampler(data, GridGenerator(loc, "affine", target_shape))


def Correlation(data1, data2, kernel_size=1, max_displacement=1, stride1=1, stride2=1, pad_size=0, is_multiply=True):
    """FlowNet correlation layer (``correlation-inl.h``): for every displacement (multiples of ``stride2`` up to ``max_displacement``) the patch
    product (or absolute difference) averaged over ``kernel_size² · C``.  Output ``[B, D², Ho, Wo]``."""
    a = TF.pad(_t(data1), (pad_size,) * 4); b = TF.pad(_t(data2), (pad_size,) * 4)
    B, C, H, W = a.shape
    kr = (kernel_size - 1) // 2
    border = max_displacement + kr
    Ho = int(_math.ceil((H - 2 * border) / stride1)); Wo = int(_math.ceil((W - 2 * border) / stride1))
    r = max_displacement // stride2
    outs = []
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            oy, ox = dy * stride2, dx * stride2
            pa = a[:, :, border - kr: H - border + kr, border - kr: W - border + kr]
            pb = b[:, :, border - kr + oy: H - border + kr + oy, border - kr + ox: W - border + kr + ox]
            prod = pa * pb if is_multiply else (pa - pb).abs()
            s = TF.avg_pool2d(prod.sum(1, keepdim=True), kernel_size, stride=1) * (kernel_size * kernel_size) if kernel_size > 1 else prod.sum(1, keepdim=True)
            outs.append(s[:, :, ::stride1, ::stride1][:, :, :Ho, :Wo] / (kernel_size * kernel_size * C))
    return _W(torch.cat(outs, 1))


def ROIPooling(data, rois, pooled_size, spatial_scale):
    from .contrib import ROIPooling as _rp
    return _rp(data, rois, pooled_size, spatial_scale)


def SVMOutput(data, label=None, margin=1.0, regularization_coefficient=1.0, use_linear=False):
    """Forward is the identity (``svm_output-inl.h``); the hinge gradient is what ``gluon.loss.HingeLoss`` / ``SquaredHingeLoss`` provide."""
    return _W(_t(data).clone())


def Crop(*data, offset=(0, 0), h_w=(0, 0), center_crop=False, num_args=1):
    """Crop ``data[0]`` spatially to ``h_w`` or to the size of ``data[1]`` (``crop-inl.h``)."""
    x = _t(data[0])
    th, tw = (int(data[1].shape[2]), int(data[1].shape[3])) if len(data) > 1 else (int(h_w[0]), int(h_w[1]))
    oy, ox = ((x.shape[2] - th) // 2, (x.shape[3] - tw) // 2) if center_crop else (int(offset[0]), int(offset[1]))
    return _W(x[:, :, oy:oy + th, ox:ox + tw])


def histogram(a, bins=10, range=None):
    """``(counts, bin_edges)`` like numpy (``tensor/histogram-inl.h``); ``bins`` may be an edge array."""
    x = _t(a).float().flatten()
    if isinstance(bins, NDArray):
        edges = _t(bins).float()
        idx = torch.bucketize(x, edges, right=True) - 1
        idx = torch.where(x == edges[-1], torch.full_like(idx, edges.numel() - 2), idx)
        ok = (idx >= 0) & (idx < edges.numel() - 1)
        return _W(torch.bincount(idx[ok], minlength=edges.numel() - 1).to(torch.int64)), _W(edges)
    lo, hi = (float(x.min()), float(x.max())) if range is None else (float(range[0]), float(range[1]))
    return _W(torch.histc(x, int(bins), lo, hi).to(torch.int64)), _W(torch.linspace(lo, hi, int(bins) + 1, device=x.device))


def ravel_multi_index(data, shape):
    idx = _t(data).long(); out = torch.zeros_like(idx[0]); mul = 1
    for d in _bi.range(len(shape) - 1, -1, -1):
        out = out + idx[d] * mul; mul *= int(shape[d])
    return _W(out.to(_t(data).dtype))


def unravel_index(data, shape):
    idx = _t(data).long(); outs = []
    for d in _bi.range(len(shape) - 1, -1, -1):
        outs.append(idx % int(shape[d])); idx = idx // int(shape[d])
    return _W(torch.stack(outs[::-1]).to(_t(data).dtype))


def choose_element_0index(lhs, rhs):
    from .ndarray import pick as _pick
    return _pick(lhs, rhs, axis=1)


def fill_element_0index(lhs, mhs, rhs):
    out = _t(lhs).clone(); out[torch.arange(out.shape[0], device=out.device), _t(rhs).long()] = _t(mhs)
    return _W(out)


crop = slice
Softmax = SoftmaxOutput


# ------------------------------------------------------------------------------------------------ optimizer update ops (optimizer_op-inl.h)
def _prep_grad(weight, grad, wd, rescale_grad, clip_gradient):
    g = _t(grad) * rescale_grad
    if clip_gradient is not None and clip_gradient >= 0:
        g = g.clamp(-clip_gradient, clip_gradient)
    return g + wd * _t(weight)


def _ret(weight, new, out):
    tgt = weight if out is None else out
    with torch.no_grad():
        _t(tgt).copy_(new)
    return tgt


def sgd_update(weight, grad, lr, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, lazy_update=True, out=None):
    """``w -= lr * (clip(rescale * g) + wd * w)`` (``optimizer_op-inl.h:86-103``); in place on ``weight`` unless ``out`` is given."""
    with torch.no_grad():
        return _ret(weight, _t(weight) - lr * _prep_grad(weight, grad, wd, rescale_grad, clip_gradient), out)


def sgd_mom_update(weight, grad, mom, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, lazy_update=True, out=None):
    with torch.no_grad():
        _t(mom).mul_(momentum).sub_(lr * _prep_grad(weight, grad, wd, rescale_grad, clip_gradient))
        return _ret(weight, _t(weight) + _t(mom), out)


def mp_sgd_update(weight, grad, weight32, lr, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, lazy_update=True, out=None):
    """Multi-precision SGD: the fp32 master ``weight32`` is updated, ``weight`` receives its cast (``optimizer_op-inl.h:359-400``)."""
    with torch.no_grad():
        g = _t(grad).float() * rescale_grad
        if clip_gradient is not None and clip_gradient >= 0:
            g = g.clamp(-clip_gradient, clip_gradient)
        _t(weight32).sub_(lr * (g + wd * _t(weight32)))
        return _ret(weight, _t(weight32).to(_t(weight).dtype), out)


def mp_sgd_mom_update(weight, grad, mom, weight32, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, lazy_update=True, out=None):
    with torch.no_grad():
        g = _t(grad).float() * rescale_grad
        if clip_gradient is not None and clip_gradient >= 0:
            g = g.clamp(-clip_gradient, clip_gradient)
        _t(mom).mul_(momentum).sub_(lr * (g + wd * _t(weight32)))
        _t(weight32).add_(_t(mom))
        return _ret(weight, _t(weight32).to(_t(weight).dtype), out)


def adam_update(weight, grad, mean, var, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0,
                lazy_update=True, out=None):
    """One Adam step WITHOUT bias correction (the Python optimizer folds it into ``lr``) (``optimizer_op-inl.h:840-873``).  On CUDA fp32
    tensors this is the single-pass native kernel (``csrc/kernels/optim.cu``)."""
    w, g, m, v = _t(weight), _t(grad), _t(mean), _t(var)
    if out is None and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and g.is_contiguous():
        from ..ops import native
        if native.available():
            from ..ops import _native_api as n
            n.adam_update(w, g, m, v, lr, beta1, beta2, epsilon, wd, rescale_grad, clip_gradient if clip_gradient is not None else -1.0)
            return weight
    with torch.no_grad():
        gg = _prep_grad(weight, grad, wd, rescale_grad, clip_gradient)
        m.mul_(beta1).add_(gg * (1 - beta1)); v.mul_(beta2).add_(gg * gg * (1 - beta2))
        return _ret(weight, w - lr * m / (v.sqrt() + epsilon), out)


def rmsprop_update(weight, grad, n, lr, gamma1=0.95, epsilon=1e-8, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, clip_weights=-1.0, out=None):
    with torch.no_grad():
        g = _prep_grad(weight, grad, wd, rescale_grad, clip_gradient)
        _t(n).mul_(gamma1).add_(g * g * (1 - gamma1))
        new = _t(weight) - lr * g / (_t(n) + epsilon).sqrt()
        if clip_weights is not None and clip_weights > 0:
            new = new.clamp(-clip_weights, clip_weights)
        return _ret(weight, new, out)


def rmspropalex_update(weight, grad, n, g, delta, lr, gamma1=0.95, gamma2=0.9, epsilon=1e-8, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0,
                       clip_weights=-1.0, out=None):
    with torch.no_grad():
        gr = _prep_grad(weight, grad, wd, rescale_grad, clip_gradient)
        _t(n).mul_(gamma1).add_(gr * gr * (1 - gamma1)); _t(g).mul_(gamma1).add_(gr * (1 - gamma1))
        _t(delta).mul_(gamma2).sub_(lr * gr / (_t(n) - _t(g) ** 2 + epsilon).sqrt())
        new = _t(weight) + _t(delta)
        if clip_weights is not None and clip_weights > 0:
            new = new.clamp(-clip_weights, clip_weights)
        return _ret(weight, new, out)


def ftrl_update(weight, grad, z, n, lr, lamda1=0.01, beta=1.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, out=None):
    with torch.no_grad():
        g = _t(grad) * rescale_grad
        if clip_gradient is not None and clip_gradient >= 0:
            g = g.clamp(-clip_gradient, clip_gradient)
        w = _t(weight)
        _t(z).add_(g - ((_t(n) + g * g).sqrt() - _t(n).sqrt()) * w / lr)
        _t(n).add_(g * g)
        new = (torch.sign(_t(z)) * lamda1 - _t(z)) / ((beta + _t(n).sqrt()) / lr + wd) * (_t(z).abs() > lamda1)
        return _ret(weight, new, out)


def signsgd_update(weight, grad, lr, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, out=None):
    with torch.no_grad():
        return _ret(weight, (1 - lr * wd) * _t(weight) - lr * torch.sign(_t(grad)), out)


def signum_update(weight, grad, mom, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, wd_lh=0.0, out=None):
    with torch.no_grad():
        g = _prep_grad(weight, grad, wd, rescale_grad, clip_gradient)
        _t(mom).mul_(momentum).sub_((1 - momentum) * g)
        return _ret(weight, (1 - lr * wd_lh) * _t(weight) + lr * torch.sign(_t(mom)), out)


def ftml_update(weight, grad, d, v, z, lr, t, beta1=0.6, beta2=0.999, epsilon=1e-8, wd=0.0, rescale_grad=1.0, clip_grad=-1.0, out=None):
    with torch.no_grad():
        g = _prep_grad(weight, grad, wd, rescale_grad, clip_grad)
        _t(v).mul_(beta2).add_(g * g * (1 - beta2))
        d_t = (1 - beta1 ** t) / lr * ((_t(v) / (1 - beta2 ** t)).sqrt() + epsilon)
        sigma = d_t - beta1 * _t(d)
        _t(z).mul_(beta1).add_((1 - beta1) * g - sigma * _t(weight))
        _t(d).copy_(d_t)
        return _ret(weight, -_t(z) / d_t, out)


# ------------------------------------------------------------------------------------------------ fused RNN op, versioned aliases, misc
def RNN(data, parameters, state, state_cell=None, state_size=None, num_layers=1, bidirectional=False, mode="lstm", p=0.0, state_outputs=False,
        projection_size=None, **kw):
    """The fused multi-layer RNN operator (``src/operator/rnn-inl.h``): ``data [T, N, C]``, ``parameters`` = ONE flat vector holding, layer by
    layer and direction by direction, all ``W_i2h, W_h2h`` blocks and then all ``b_i2h, b_h2h`` blocks (gate order LSTM i,f,g,o; GRU r,z,n),
    ``state`` / ``state_cell`` ``[L*D, N, H]``.  Returns the output sequence, plus the final states with ``state_outputs=True``."""
    x, flat = _t(data), _t(parameters).reshape(-1)
    H, L, D = int(state_size), int(num_layers), (2 if bidirectional else 1)
    G = {"rnn_relu": 1, "rnn_tanh": 1, "lstm": 4, "gru": 3}[mode]
    T, N, C = x.shape
    w_shapes = []
    for layer in _bi.range(L):
        cin = C if layer == 0 else D * H
        for _ in _bi.range(D):
            w_shapes += [(G * H, cin), (G * H, H)]
    pos, weights = 0, []
    for shp in w_shapes:
        n = shp[0] * shp[1]
        weights.append(flat[pos:pos + n].view(shp)); pos += n
    biases = []
    for _ in w_shapes:
        biases.append(flat[pos:pos + G * H]); pos += G * H
    if pos != flat.numel():
        raise ValueError("RNN: parameter vector has %d elements, the configuration needs %d" % (flat.numel(), pos))
    h0 = _t(state); c0 = _t(state_cell) if state_cell is not None else None
    seq, hN, cN = x, [], []
    for layer in _bi.range(L):
        outs_dir = []
        for d in _bi.range(D):
            k = layer * D + d
            wi, wh, bi, bh = weights[2 * k], weights[2 * k + 1], biases[2 * k], biases[2 * k + 1]
            h = h0[k]; c = c0[k] if c0 is not None else None
            steps = _bi.range(T) if d == 0 else _bi.range(T - 1, -1, -1)
            outs = [None] * T
            xi_all = TF.linear(seq, wi, bi)                                    # input projections of every step in one GEMM
            for t in steps:
                gh = TF.linear(h, wh, bh)
                if mode == "lstm":
                    i, f, g, o = (xi_all[t] + gh).chunk(4, dim=-1)
                    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                    h = torch.sigmoid(o) * torch.tanh(c)
                elif mode == "gru":
                    xr, xz, xn = xi_all[t].chunk(3, dim=-1); hr, hz, hn = gh.chunk(3, dim=-1)
                    r = torch.sigmoid(xr + hr); z = torch.sigmoid(xz + hz)
                    n_ = torch.tanh(xn + r * hn)
                    h = (1 - z) * n_ + z * h
                else:
                    h = (torch.relu if mode == "rnn_relu" else torch.tanh)(xi_all[t] + gh)
                outs[t] = h
            outs_dir.append(torch.stack(outs, 0)); hN.append(h)
            if c is not None:
                cN.append(c)
        seq = torch.cat(outs_dir, dim=-1) if D == 2 else outs_dir[0]
        if p > 0 and layer + 1 < L:
            from .. import autograd as _ag
            seq = TF.dropout(seq, p, _ag.is_training())
    if not state_outputs:
        return _W(seq)
    res = [_W(seq), _W(torch.stack(hN, 0))]
    if mode == "lstm":
        res.append(_W(torch.stack(cN, 0)))
    return res


def BatchNorm_v1(*a, **k): return BatchNorm(*a, **k)
def CuDNNBatchNorm(*a, **k): return BatchNorm(*a, **k)
def Convolution_v1(*a, **k): return Convolution(*a, **k)
def Pooling_v1(*a, **k): return Pooling(*a, **k)


def IdentityAttachKLSparseReg(data, sparseness_target=0.1, penalty=0.001, momentum=0.9):
    """Identity in the forward pass; the reference attaches a KL sparsity penalty to the gradient of sigmoid activations
    (``identity_attach_KL_sparse_reg-inl.h``): ``grad += penalty * (-rho/rho_hat + (1-rho)/(1-rho_hat))`` with ``rho_hat`` the batch mean."""
    x = _t(data)

    class _F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            ctx.save_for_backward(t)
            return t.view_as(t)

        @staticmethod
        def backward(ctx, g):
            (t,) = ctx.saved_tensors
            rho_hat = t.mean(0, keepdim=True).clamp(1e-6, 1 - 1e-6)
            return g + penalty * (-sparseness_target / rho_hat + (1 - sparseness_target) / (1 - rho_hat))
    return _W(_F.apply(x))


def cast_storage(data, stype):
    from .sparse import cast_storage as _cs
    return _cs(data, stype)


def linalg_gelqf(A):
    """LQ factorisation ``A = L Q`` with ``Q`` having orthonormal rows (``la_op.h`` gelqf); returns ``(Q, L)`` like the reference."""
    q, r = torch.linalg.qr(_t(A).transpose(-1, -2), mode="reduced")
    # make the diagonal of L non-negative (LAPACK convention used by the reference's tests)
    sgn = torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)); sgn = torch.where(sgn == 0, torch.ones_like(sgn), sgn)
    q = q * sgn.unsqueeze(-2); r = r * sgn.unsqueeze(-1)
    return _W(q.transpose(-1, -2)), _W(r.transpose(-1, -2))


def linalg_syevd(A):
    """Symmetric eigendecomposition ``A = U^T diag(L) U`` (rows of ``U`` are eigenvectors, eigenvalues ascending); returns ``(U, L)``."""
    w, v = torch.linalg.eigh(_t(A))
    return _W(v.transpose(-1, -2)), _W(w)

__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "TF", "OF", "NDArray", "torch_dtype", "annotations")]
