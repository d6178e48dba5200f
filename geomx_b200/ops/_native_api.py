"""Python wrappers over the C ABI of ``libgeomx_kernels.so``.  All functions take raw CUDA ``torch.Tensor`` s, launch on
the current stream and return nothing (or the output tensor); shapes/dtypes are asserted here, not in the kernels."""
from __future__ import annotations

import ctypes
import math

import torch

from . import native as _n

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _lib():
    return _n.require()


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _ck(rc, what):
    _n.launch_count += 1
    if rc != 0:
        raise RuntimeError("%s failed (rc=%d)" % (what, rc))


def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expected contiguous CUDA fp32 tensor"
    return t


# --------------------------------------------------------------------------------------------------------------- GEMM
def set_gemm_precision(mode):
    """Process-wide multiply precision of the tcgen05 GEMM: ``"3xtf32"`` (default; fp32-accurate: hi/lo split of both operands, three MMAs per
    K step, fp32 TMEM accumulation — matches the reference's fp32 SGEMM, linalg_impl.h:196-214) or ``"tf32"`` (one MMA, 10-bit mantissa)."""
    return _lib().gx_gemm_set_precision(1 if str(mode).lower() in ("tf32", "1", "fast") else 3)


def gemm_precision():
    return "tf32" if _lib().gx_gemm_get_precision() == 1 else "3xtf32"


def gemm(A, B, D, a_mn=False, b_mn=False, bias=None, mask=None, colsum=None, relu=False, accumulate=False,
         store_nchw_hw=0, alpha=1.0, split_k=1, M=None, N=None, K=None, lda=None, ldb=None, ldd=None, force_simt=False):
    """D[M,N] = epi(alpha * A·Bᵀ).  A is [M,K] (K-major) or [K,M] (MN-major, ``a_mn``); B is [N,K] or [K,N] (``b_mn``)."""
    if M is None:
        M = A.shape[1] if a_mn else A.shape[0]
    if K is None:
        K = A.shape[0] if a_mn else A.shape[1]
    if N is None:
        N = B.shape[1] if b_mn else B.shape[0]
    lda = A.stride(0) if lda is None else lda
    ldb = B.stride(0) if ldb is None else ldb
    if ldd is None:
        ldd = D.stride(0) if (store_nchw_hw == 0 and D.dim() >= 2) else N
    ldmask = mask.stride(0) if mask is not None else 0
    lib = _lib()
    rc = -1
    if not force_simt:
        rc = lib.gx_gemm_tf32(_p(A), lda, int(a_mn), _p(B), ldb, int(b_mn), M, N, K, _p(D), ldd, _p(bias), _p(mask), ldmask,
                              _p(colsum), int(relu), int(accumulate), 1 if store_nchw_hw else 0, int(store_nchw_hw), float(alpha),
                              int(split_k), _s())
    if rc == -1:  # TMA alignment not satisfied -> CUDA-core fallback kernel (same contract)
        rc = lib.gx_gemm_simt(_p(A), lda, int(a_mn), _p(B), ldb, int(b_mn), M, N, K, _p(D), ldd, _p(bias), _p(mask), ldmask,
                              _p(colsum), int(relu), int(accumulate), 1 if store_nchw_hw else 0, int(store_nchw_hw), float(alpha), _s())
    _ck(rc, "gemm")
    return D


def gemm_pool(A, B, pooled, idx, OH, OW, bias=None, relu=True, a_mn=False, b_mn=False, alpha=1.0):
    """Convolution-as-GEMM with bias + ReLU + 2x2/2 max-pool fused into the tcgen05 epilogue.  ``A``: [images*OH*OW, K] im2col rows,
    ``B``: [N, K] filters; writes ``pooled`` [images, N, OH/2, OW/2] and ``idx`` (uint8 arg-max).  Returns False when the geometry does not
    fit the in-warp pooling (caller falls back to gemm + maxpool2x2_fwd)."""
    M = A.shape[1] if a_mn else A.shape[0]
    K = A.shape[0] if a_mn else A.shape[1]
    N = B.shape[1] if b_mn else B.shape[0]
    rc = _lib().gx_gemm_tf32_pool(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), M, N, K, _p(pooled), _p(bias), int(relu), OH * OW, OW,
                                  _p(idx), float(alpha), _s())
    if rc == -1:
        return False
    _ck(rc, "gemm_pool")
    return True


def mlp_chain(x, w0, b0, w1, b1, w2, b2, label, loss, logits, dw0, db0, dw1, db1, dw2, db2, dx):
    """Dense(D1,relu) -> Dense(D2,relu) -> Dense(C) -> softmax-CE, forward and backward, in ONE cluster launch (csrc/kernels/mlp_chain.cu).
    Returns False when the shapes are not the compiled 512 -> 256 -> 128 -> C<=16, batch <= 32 (callers then use the per-layer kernels)."""
    B, D0 = x.shape
    D1, D2, C = w0.shape[0], w1.shape[0], w2.shape[0]
    rc = _lib().gx_mlp_chain_fwd_bwd(_p(x), _p(w0), _p(b0), _p(w1), _p(b1), _p(w2), _p(b2), _p(label), _p(loss), _p(logits), _p(dw0), _p(db0),
                                     _p(dw1), _p(db1), _p(dw2), _p(db2), _p(dx), B, D0, D1, D2, C, _s())
    if rc == -1:
        return False
    _ck(rc, "mlp_chain")
    return True


# --------------------------------------------------------------------------------------------------------------- demo-CNN direct convolutions
def cnn_fwd(x, w0, b0, w1, b1, a1, idx1, a2, idx2, carry=None, x_keep=None):
    """Conv(16,k5)+ReLU+MaxPool2 -> Conv(32,k5)+ReLU+MaxPool2 of the demo CNN in one launch (csrc/kernels/cnn_direct.cu), fp32 FMA.
    ``carry=(src, dst)``: also copy the small fp32 vector ``src`` to ``dst`` (the labels of a look-ahead step, see models/cnn.py);
    ``x_keep``: also store a copy of the image batch there (its backward pass runs after the next batch has arrived in ``x``)."""
    src, dst = carry if carry is not None else (None, None)
    _ck(_lib().gx_cnn_fwd(_p(x), _p(w0), _p(b0), _p(w1), _p(b1), _p(a1), _p(idx1), _p(a2), _p(idx2), _p(src), _p(dst),
                          0 if src is None else src.numel(), _p(x_keep), x.shape[0], _s()), "cnn_fwd")


def cnn_bwd(x, w1, a1, idx1, a2, idx2, da2, dw0, db0):
    """conv1 data gradient + pool/ReLU backward of both layers + conv0 weight / bias gradient (accumulated with atomics into dw0 / db0)."""
    _ck(_lib().gx_cnn_bwd(_p(x), _p(w1), _p(a1), _p(idx1), _p(a2), _p(idx2), _p(da2), _p(dw0), _p(db0), x.shape[0], _s()), "cnn_bwd")


def cnn_bwd_all(x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dw1, db1):
    """The whole convolution backward pass (``cnn_bwd`` + ``cnn_wgrad1``) as one heterogeneous-grid launch."""
    _ck(_lib().gx_cnn_bwd_all(_p(x), _p(w1), _p(a1), _p(idx1), _p(a2), _p(idx2), _p(da2), _p(dw0), _p(db0), _p(dw1), _p(db1), x.shape[0], _s()),
        "cnn_bwd_all")


def cnn_bwd_exchange(x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dw1, db1, params, tile_list, n_active):
    """``cnn_bwd_all`` whose last CTAs also perform the exchange of the conv keys (``params``: the channel's FabricParams block from
    ``HipsFabric.channel_fused_args``): convolution backward + push + server tiers + optimizer + pull in one launch."""
    import ctypes
    _ck(_lib().gx_cnn_bwd_exchange(_p(x), _p(w1), _p(a1), _p(idx1), _p(a2), _p(idx2), _p(da2), _p(dw0), _p(db0), _p(dw1), _p(db1), x.shape[0],
                                   ctypes.byref(params), _p(tile_list), int(n_active), _s()), "cnn_bwd_exchange")


def cnn_wgrad1(a1, a2, idx2, da2, dw1, db1):
    """conv1 weight / bias gradient (overwrites dw1 / db1).  Returns False for batch sizes the kernel does not cover (odd or > 64)."""
    rc = _lib().gx_cnn_wgrad1(_p(a1), _p(a2), _p(idx2), _p(da2), _p(dw1), _p(db1), a1.shape[0], _s())
    if rc == -1:
        return False
    _ck(rc, "cnn_wgrad1")
    return True


# --------------------------------------------------------------------------------------------------------------- depthwise convolution
def depthwise_fwd(x, w, bias, y, stride, padding, relu=False):
    N, C, H, W = x.shape
    rc = _lib().gx_depthwise_fwd(_p(x), _p(w), _p(bias), _p(y), N, C, H, W, w.shape[2], w.shape[3], stride[0], stride[1], padding[0], padding[1],
                                 int(relu), _s())
    if rc == -1:
        return False
    _ck(rc, "depthwise_fwd")
    return True


def depthwise_dgrad(dy, w, dx, stride, padding):
    N, C, H, W = dx.shape
    _ck(_lib().gx_depthwise_dgrad(_p(dy), _p(w), _p(dx), N, C, H, W, w.shape[2], w.shape[3], stride[0], stride[1], padding[0], padding[1], _s()),
        "depthwise_dgrad")


def depthwise_wgrad(x, dy, dw, db, stride, padding):
    N, C, H, W = x.shape
    _ck(_lib().gx_depthwise_wgrad(_p(x), _p(dy), _p(dw), _p(db), N, C, H, W, dw.shape[2], dw.shape[3], stride[0], stride[1], padding[0], padding[1],
                                  _s()), "depthwise_wgrad")


# --------------------------------------------------------------------------------------------------------------- conv / pool
def conv_out_hw(H, W, KH, KW, sh, sw, ph, pw):
    return (H + 2 * ph - KH) // sh + 1, (W + 2 * pw - KW) // sw + 1


def conv_supported(x, w, stride, padding):
    return x.dim() == 4 and w.dim() == 4 and x.shape[1] == w.shape[1]


def im2col(x, KH, KW, stride=(1, 1), padding=(0, 0), out=None):
    N, C, H, W = x.shape
    OH, OW = conv_out_hw(H, W, KH, KW, stride[0], stride[1], padding[0], padding[1])
    K = C * KH * KW
    ldc = (K + 3) // 4 * 4
    if out is None:
        out = torch.empty(N * OH * OW, ldc, dtype=torch.float32, device=x.device)
    _ck(_lib().gx_im2col(_p(_f32c(x)), _p(out), N, C, H, W, KH, KW, stride[0], stride[1], padding[0], padding[1], ldc, _s()), "im2col")
    return out


def col2im(dcol, x_shape, KH, KW, stride=(1, 1), padding=(0, 0), out=None):
    N, C, H, W = x_shape
    if out is None:
        out = torch.empty(x_shape, dtype=torch.float32, device=dcol.device)
    _ck(_lib().gx_col2im(_p(dcol), _p(out), N, C, H, W, KH, KW, stride[0], stride[1], padding[0], padding[1], dcol.stride(0), _s()), "col2im")
    return out


def nchw_to_rows(x, out=None):
    N, C = x.shape[0], x.shape[1]
    HW = x.numel() // (N * C)
    if out is None:
        out = torch.empty(N * HW, C, dtype=torch.float32, device=x.device)
    _ck(_lib().gx_nchw_to_rows(_p(_f32c(x)), _p(out), N, C, HW, _s()), "nchw_to_rows")
    return out


def colsum(x, out=None, accumulate=False):
    R, Cc = x.shape
    if out is None:
        out = torch.empty(Cc, dtype=torch.float32, device=x.device)
    _ck(_lib().gx_colsum(_p(x), _p(out), R, Cc, x.stride(0), int(accumulate), _s()), "colsum")
    return out


def chansum_nchw(x, out=None):
    N, C = x.shape[0], x.shape[1]
    HW = x.numel() // (N * C)
    if out is None:
        out = torch.empty(C, dtype=torch.float32, device=x.device)
    _ck(_lib().gx_chansum_nchw(_p(_f32c(x)), _p(out), N, C, HW, _s()), "chansum")
    return out


def relu_fwd(x, out=None):
    out = torch.empty_like(x) if out is None else out
    _ck(_lib().gx_relu_fwd(_p(_f32c(x)), _p(out), x.numel(), _s()), "relu_fwd")
    return out


def relu_bwd(y, dy, out=None):
    out = torch.empty_like(dy) if out is None else out
    _ck(_lib().gx_relu_bwd(_p(_f32c(y)), _p(_f32c(dy)), _p(out), y.numel(), _s()), "relu_bwd")
    return out


def maxpool2x2_fwd(x, out=None, idx=None):
    N, C, H, W = x.shape
    if out is None:
        out = torch.empty(N, C, H // 2, W // 2, dtype=torch.float32, device=x.device)
    if idx is None:
        idx = torch.empty(N, C, H // 2, W // 2, dtype=torch.uint8, device=x.device)
    _ck(_lib().gx_maxpool2x2_fwd(_p(_f32c(x)), _p(out), _p(idx), N * C, H, W, _s()), "maxpool_fwd")
    return out, idx


def maxpool2x2_bwd(dy, idx, x_shape, out=None):
    N, C, H, W = x_shape
    if out is None:
        out = torch.empty(x_shape, dtype=torch.float32, device=dy.device)
    _ck(_lib().gx_maxpool2x2_bwd(_p(_f32c(dy)), _p(idx), _p(out), N * C, H, W, _s()), "maxpool_bwd")
    return out


def pool_relu_bwd_rows(dpooled, pooled, idx, dz_rows, dbias):
    N, C, PH, PW = pooled.shape
    _ck(_lib().gx_pool_relu_bwd_rows(_p(dpooled), _p(pooled), _p(idx), _p(dz_rows), _p(dbias), N, C, PH, PW, _s()), "pool_relu_bwd_rows")


def conv_relu_pool_fwd(x, w, b, y, idx):
    N, Cin, H, W = x.shape
    Cout, _, KH, KW = w.shape
    _ck(_lib().gx_conv_relu_pool_fwd(_p(x), _p(w), _p(b), _p(y), _p(idx), N, Cin, H, W, Cout, KH, KW, _s()), "conv_relu_pool_fwd")


def conv_relu_pool_wgrad(x, dpooled, pooled, idx, dw, db, w_shape):
    N, Cin, H, W = x.shape
    Cout, _, KH, KW = w_shape
    _ck(_lib().gx_conv_relu_pool_wgrad(_p(x), _p(dpooled), _p(pooled), _p(idx), _p(dw), _p(db), N, Cin, H, W, Cout, KH, KW, _s()),
        "conv_relu_pool_wgrad")


def conv_relu_pool_im2col_fwd(x, w, b, y, idx, col, KH2, KW2):
    N, Cin, H, W = x.shape
    Cout, _, KH, KW = w.shape
    _ck(_lib().gx_conv_relu_pool_im2col_fwd(_p(x), _p(w), _p(b), _p(y), _p(idx), _p(col), N, Cin, H, W, Cout, KH, KW, KH2, KW2, col.stride(0), _s()),
        "conv_relu_pool_im2col_fwd")


def conv_relu_pool_wgrad_col2im(x, dcol, pooled, idx, dw, db, w_shape, KH2, KW2):
    N, Cin, H, W = x.shape
    Cout, _, KH, KW = w_shape
    _ck(_lib().gx_conv_relu_pool_wgrad_col2im(_p(x), _p(dcol), _p(pooled), _p(idx), _p(dw), _p(db), N, Cin, H, W, Cout, KH, KW, KH2, KW2, dcol.stride(0), _s()),
        "conv_relu_pool_wgrad_col2im")


# --------------------------------------------------------------------------------------------------------------- loss / head
def softmax_ce_fwd(logits, label, out=None):
    R, Cc = logits.shape
    out = torch.empty(R, dtype=torch.float32, device=logits.device) if out is None else out
    _ck(_lib().gx_softmax_ce_fwd(_p(_f32c(logits)), _p(_f32c(label)), _p(out), R, Cc, _s()), "softmax_ce_fwd")
    return out


def softmax_ce_bwd(logits, label, dloss, out=None):
    R, Cc = logits.shape
    out = torch.empty_like(logits) if out is None else out
    _ck(_lib().gx_softmax_ce_bwd(_p(logits), _p(label), _p(dloss), _p(out), R, Cc, _s()), "softmax_ce_bwd")
    return out


def head_fwd_bwd(a, W, bias, label, loss, logits, dW, db, da, dbias_prev, relu_mask=True):
    B, K = a.shape
    Cc = W.shape[0]
    _ck(_lib().gx_head_fwd_bwd(_p(a), _p(W), _p(bias), _p(label), _p(loss), _p(logits), _p(dW), _p(db), _p(da), _p(dbias_prev),
                               B, K, Cc, int(relu_mask), _s()), "head_fwd_bwd")


# --------------------------------------------------------------------------------------------------------------- optimizers
_KIND = {"sgd": 0, "adam": 1, "dcasgd": 2}


def arena_opt(kind, w, g, s0, s1, n, tile_mult, lr, wd, rescale, clip, momentum, b1, b2, eps, lamda, step_state, g_zero=None):
    _ck(_lib().gx_arena_opt(_KIND[kind], _p(w), _p(g), _p(s0), _p(s1), n, _p(tile_mult), lr, wd, rescale, clip, momentum, b1, b2, eps,
                            lamda, _p(step_state), _p(g_zero), _s()), "arena_opt")


def single_opt(kind, w, g, s0, s1, lr_t, wd, rescale, clip, momentum=0.0, b1=0.9, b2=0.999, eps=1e-8, lamda=0.04):
    _ck(_lib().gx_single_opt(_KIND[kind], _p(w), _p(g), _p(s0), _p(s1), w.numel(), lr_t, wd, rescale, clip, momentum, b1, b2, eps, lamda, _s()),
        "single_opt")


def adam_update(w, g, m, v, lr_t, b1, b2, eps, wd, rescale, clip):
    single_opt("adam", w, g.contiguous(), m, v, lr_t, wd, rescale, clip, 0.0, b1, b2, eps)


class _TensorEntry(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("g", ctypes.c_void_p), ("s0", ctypes.c_void_p), ("s1", ctypes.c_void_p),
                ("n", ctypes.c_longlong), ("lr_mult", ctypes.c_float), ("wd_mult", ctypes.c_float)]


_mt_cache = {}


def multi_tensor_update(optimizer, updater, idxs, ws, gs):
    """One launch for the whole parameter list (Trainer fused path).  Returns False if this optimizer has no native spec."""
    from ..ndarray import NDArray
    spec = optimizer.spec()
    kind = spec["name"]
    if kind == "sgd" and spec.get("momentum", 0.0) == 0.0:
        need = 0
    else:
        need = 1 if kind == "sgd" else 2
    for i, w in zip(idxs, ws):
        if i not in updater.states:
            updater.states[i] = optimizer.create_state_multi_precision(i, NDArray(w))
            updater.states_synced[i] = True
    optimizer._update_count(idxs)
    t = optimizer._index_update_count[idxs[0]]
    lr = optimizer.learning_rate
    lr_t = lr * math.sqrt(1.0 - spec["beta2"] ** t) / (1.0 - spec["beta1"] ** t) if kind == "adam" else lr
    key = (id(updater), tuple(w.data_ptr() for w in ws), tuple(g.data_ptr() for g in gs))
    ent = _mt_cache.get(key)
    if ent is None:
        arr = (_TensorEntry * len(ws))()
        for j, (i, w, g) in enumerate(zip(idxs, ws, gs)):
            st = updater.states[i]
            s0 = s1 = None
            if need == 1:
                s0 = st._t
            elif need == 2:
                s0, s1 = st[0]._t, st[1]._t
            p = optimizer.param_dict.get(i)
            arr[j] = _TensorEntry(w.data_ptr(), g.data_ptr(), 0 if s0 is None else s0.data_ptr(), 0 if s1 is None else s1.data_ptr(),
                                  w.numel(), p.lr_mult if p is not None else 1.0, p.wd_mult if p is not None else 1.0)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = host.to(ws[0].device)
        ent = (dev, max(w.numel() for w in ws), len(ws))
        if len(_mt_cache) > 64:
            _mt_cache.clear()
        _mt_cache[key] = ent
    dev, max_n, cnt = ent
    clip = -1.0 if optimizer.clip_gradient is None else float(optimizer.clip_gradient)
    _ck(_lib().gx_multi_tensor_opt(_KIND[kind], _p(dev), cnt, max_n, lr_t, float(optimizer.wd), float(optimizer.rescale_grad), clip,
                                   float(spec.get("momentum", 0.0)), float(spec.get("beta1", 0.9)), float(spec.get("beta2", 0.999)),
                                   float(spec.get("epsilon", 1e-8)), float(spec.get("lamda", 0.04)), _s()), "multi_tensor_opt")
    return True


def nary_sum(out, parts):
    assert 1 <= len(parts) <= 8
    arr = (ctypes.c_void_p * len(parts))(*[p.data_ptr() for p in parts])
    _ck(_lib().gx_nary_sum(_p(out), arr, len(parts), out.numel(), _s()), "nary_sum")
    return out


def scale_cast(x, out_dtype, scale=1.0, out=None):
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device) if out is None else out
    _ck(_lib().gx_scale_cast(_p(x.contiguous()), _DT[x.dtype], _p(out), _DT[out.dtype], float(scale), x.numel(), _s()), "scale_cast")
    return out


# --------------------------------------------------------------------------------------------------------------- compression
def quantize_2bit(grad, residual, out, thr):
    _ck(_lib().gx_quantize_2bit(_p(_f32c(grad)), _p(residual), _p(out), grad.numel(), thr, _s()), "quantize_2bit")


def dequantize_2bit(packed, out, thr, accumulate=False):
    _ck(_lib().gx_dequantize_2bit(_p(packed), _p(out), out.numel(), thr, int(accumulate), _s()), "dequantize_2bit")


class _BscSeg(ctypes.Structure):
    _fields_ = [("grad", ctypes.c_void_p), ("u", ctypes.c_void_p), ("v", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("n", ctypes.c_longlong), ("k", ctypes.c_int), ("sample", ctypes.c_int), ("k_sample", ctypes.c_int)]


def _segs_to_dev(segs, device):
    arr = (_BscSeg * len(segs))(*segs)
    assert ctypes.sizeof(_BscSeg) == _lib().gx_bsc_seg_size()
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


def bsc_segment(grad, u, v, out, k, sample, k_sample):
    return _BscSeg(grad.data_ptr() if grad is not None else 0, 0 if u is None else u.data_ptr(), 0 if v is None else v.data_ptr(),
                   out.data_ptr(), (grad if grad is not None else v).numel(), k, sample, k_sample)


def bsc_compress_batch(seg_table_dev, num, max_sample, momentum=0.9):
    _ck(_lib().gx_bsc_compress_batch(_p(seg_table_dev), num, max_sample, momentum, _s()), "bsc_compress")


def bsc_compress(grad, u, v, out, k, sample, k_sample, momentum=0.9):
    tab = _segs_to_dev([bsc_segment(_f32c(grad), u, v, out, k, sample, k_sample)], grad.device)
    bsc_compress_batch(tab, 1, sample, momentum)
    return out


def bsc_pull_compress(dense, out, k):
    tab = _segs_to_dev([_BscSeg(dense.data_ptr(), 0, 0, out.data_ptr(), dense.numel(), k, 0, 0)], dense.device)
    _ck(_lib().gx_bsc_pull_compress_batch(_p(tab), 1, _s()), "bsc_pull_compress")
    return out


def bsc_decompress(zipped, out, accumulate=False):
    _ck(_lib().gx_bsc_decompress(_p(zipped), _p(out), out.numel(), zipped.numel() // 2, int(accumulate), _s()), "bsc_decompress")
    return out


def fp8_block_quantize(x, residual, q, scale):
    _ck(_lib().gx_fp8_block_quantize(_p(_f32c(x)), _p(residual), _p(q), _p(scale), x.numel(), _s()), "fp8_quantize")


def fp8_block_dequantize(q, scale, out, accumulate=False):
    _ck(_lib().gx_fp8_block_dequantize(_p(q), _p(scale), _p(out), out.numel(), int(accumulate), _s()), "fp8_dequantize")


def dgt_contrib(g, contrib, block_elems, alpha, first):
    _ck(_lib().gx_dgt_contrib(_p(g), _p(contrib), g.numel(), block_elems, alpha, int(first), _s()), "dgt_contrib")


# --------------------------------------------------------------------------------------------------------------- row-sparse helpers
def unique_i64(ids):
    """Sorted unique of a CUDA int64 vector (CUB radix sort + select; the count is read back, as UniqueImplGPU does)."""
    ids = ids.reshape(-1).contiguous()
    n = ids.numel()
    if n == 0:
        return ids
    lib = _lib()
    ws = torch.empty(int(lib.gx_unique_i64_workspace(n)) + 16, dtype=torch.uint8, device=ids.device)
    tmp, out = torch.empty_like(ids), torch.empty_like(ids)
    cnt = torch.zeros(1, dtype=torch.int32, device=ids.device)
    _ck(lib.gx_unique_i64(_p(ids), n, _p(tmp), _p(out), _p(cnt), _p(ws), ws.numel(), _s()), "unique_i64")
    return out[:int(cnt.item())]


def gather_rows(src, ids, out=None):
    L = src.numel() // src.shape[0]
    out = torch.empty((ids.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device) if out is None else out
    _ck(_lib().gx_gather_rows(_p(_f32c(src)), _p(ids.contiguous()), _p(out), ids.numel(), L, _s()), "gather_rows")
    return out


def scatter_rows(dst, ids, rows, add=False):
    L = dst.numel() // dst.shape[0]
    _ck(_lib().gx_scatter_rows(_p(_f32c(dst)), _p(ids.contiguous()), _p(_f32c(rows)), ids.numel(), L, int(add), _s()), "scatter_rows")
    return dst


# --------------------------------------------------------------------------------------------------------------- batch norm
def bn_fwd(x, gamma, beta, rm, rv, training, momentum, eps):
    N, C = x.shape[0], x.shape[1]
    HW = x.numel() // (N * C)
    y = torch.empty_like(x)
    sm = torch.empty(C, dtype=torch.float32, device=x.device)
    si = torch.empty(C, dtype=torch.float32, device=x.device)
    _ck(_lib().gx_bn_fwd(_p(_f32c(x)), _p(gamma), _p(beta), _p(rm), _p(rv), _p(y), _p(sm), _p(si), N, C, HW, int(training), momentum, eps, _s()),
        "bn_fwd")
    return y, sm, si


def bn_bwd(x, dy, gamma, sm, si):
    N, C = x.shape[0], x.shape[1]
    HW = x.numel() // (N * C)
    dx = torch.empty_like(x)
    dg = torch.empty(C, dtype=torch.float32, device=x.device)
    dbt = torch.empty(C, dtype=torch.float32, device=x.device)
    _ck(_lib().gx_bn_bwd(_p(x), _p(_f32c(dy)), _p(gamma), _p(sm), _p(si), _p(dx), _p(dg), _p(dbt), N, C, HW, _s()), "bn_bwd")
    return dx, dg, dbt
