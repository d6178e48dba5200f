"""``mx.nd.contrib`` — the contrib operator family as imperative functions.

Parity (``src/operator/contrib/`` and ``python/mxnet/ndarray/contrib.py`` of the reference): bounding-box ops (``box_iou``, ``box_nms``,
``bipartite_matching`` — ``bounding_box-inl.h``), SSD ops (``MultiBoxPrior/Target/Detection`` — ``multibox_*.cc``), region ops (``ROIAlign``
``roi_align.cc``, ``PSROIPooling``, ``Proposal``/``MultiProposal`` ``proposal.cc``), ``DeformableConvolution``, ``fft``/``ifft``,
``count_sketch``, ``quantize``/``dequantize``/``requantize`` (``quantization/*.cc``), ``boolean_mask``, ``index_array``, ``getnnz``,
``gradientmultiplier``, control flow (``foreach``, ``while_loop``, ``cond`` — ``ndarray/contrib.py:100-460``), ``isnan/isinf/isfinite`` and the
ops that also live at top level (``AdaptiveAvgPooling2D``, ``BilinearResize2D``, ``index_copy``, ``quadratic``, ``div_sqrt_dim``,
``arange_like``, ``ctc_loss``, ``SyncBatchNorm``).

Everything is a PyTorch expression on the backing tensors (differentiable where the reference op is), box/region ops use torchvision's
kernels where they exist.  None of this is on the HiPS hot path; it is API completeness for scripts that are moved over."""
from __future__ import annotations

import math as _math

import torch
import torch.nn.functional as TF

from .ndarray import NDArray
from .op_lib import (AdaptiveAvgPooling2D, BilinearResize2D, CTCLoss, arange_like, ctc_loss, div_sqrt_dim, index_copy,  # noqa: F401
                     quadratic)

_W = NDArray


def _t(x):
    return x._t if isinstance(x, NDArray) else torch.as_tensor(x)


# ------------------------------------------------------------------------------------------------ boxes
def _corner(b, fmt):
    if fmt == "corner":
        return b
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], -1)


def _center(b):
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def _iou(a, b):
    """a [..., N, 4], b [..., M, 4] (corner) -> [..., N, M]"""
    lt = torch.maximum(a[..., :, None, :2], b[..., None, :, :2])
    rb = torch.minimum(a[..., :, None, 2:], b[..., None, :, 2:])
    wh = (rb - lt).clamp_min(0)
    inter = wh[..., 0] * wh[..., 1]
    aa = ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1]))[..., :, None]
    ab = ((b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]))[..., None, :]
    union = aa + ab - inter
    return torch.where(union > 0, inter / union.clamp_min(1e-30), torch.zeros_like(inter))


def box_iou(lhs, rhs, format="corner"):
    """IoU of every lhs box with every rhs box: out shape = lhs.shape[:-1] + rhs.shape[:-1] (``bounding_box-inl.h:560-640``)."""
    a, b = _corner(_t(lhs).float(), format), _corner(_t(rhs).float(), format)
    out = _iou(a.reshape(-1, 4), b.reshape(-1, 4))
    return _W(out.reshape(tuple(a.shape[:-1]) + tuple(b.shape[:-1])))


def _nms_keep(boxes, scores, ids, thresh, force):
    """Greedy NMS over boxes already sorted by descending score; returns a bool keep mask."""
    n = boxes.shape[0]
    keep = torch.ones(n, dtype=torch.bool, device=boxes.device)
    if n == 0:
        return keep
    iou = _iou(boxes, boxes)
    same = torch.ones_like(iou, dtype=torch.bool) if force or ids is None else ids[:, None] == ids[None, :]
    sup = (iou > thresh) & same
    for i in range(n):
        if keep[i]:
            keep &= ~(sup[i] & (torch.arange(n, device=boxes.device) > i))
    return keep


def box_nms(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1, id_index=-1, background_id=-1,
            force_suppress=False, in_format="corner", out_format="corner"):
    """Non-maximum suppression (``bounding_box-inl.h:300-520``): per batch, entries sorted by descending score; suppressed / invalid entries
    become rows of -1 at the end.  ``data`` is ``[..., N, K]``."""
    d = _t(data).float()
    shape = d.shape
    d3 = d.reshape(-1, shape[-2], shape[-1])
    out = torch.full_like(d3, -1.0)
    for b in range(d3.shape[0]):
        x = d3[b]
        scores = x[:, score_index]
        valid = scores > valid_thresh
        if id_index >= 0 and background_id >= 0:
            valid &= x[:, id_index] != background_id
        idx = torch.nonzero(valid).flatten()
        if idx.numel() == 0:
            continue
        order = idx[torch.argsort(scores[idx], descending=True, stable=True)]
        if topk > 0:
            order = order[:topk]
        xs = x[order]
        boxes = _corner(xs[:, coord_start:coord_start + 4], in_format)
        ids = xs[:, id_index] if id_index >= 0 else None
        keep = _nms_keep(boxes, xs[:, score_index], ids, overlap_thresh, force_suppress)
        kept = xs[keep].clone()
        if out_format != in_format:
            kb = _corner(kept[:, coord_start:coord_start + 4], in_format)
            kept[:, coord_start:coord_start + 4] = kb if out_format == "corner" else _center(kb)
        out[b, :kept.shape[0]] = kept
    return _W(out.reshape(shape))


box_non_maximum_suppression = box_nms


def bipartite_matching(data, is_ascend=False, threshold=None, topk=-1):
    """Greedy bipartite matching on a score matrix ``[..., N, M]`` (``bounding_box-inl.h:650-760``): repeatedly take the best remaining
    (row, col) pair whose score passes ``threshold``.  Returns (row -> col or -1, col -> row or -1)."""
    if threshold is None:
        raise ValueError("bipartite_matching requires threshold")
    d = _t(data).float()
    shape = d.shape
    d3 = d.reshape(-1, shape[-2], shape[-1])
    rows = torch.full(d3.shape[:2], -1.0, device=d.device)
    cols = torch.full((d3.shape[0], d3.shape[2]), -1.0, device=d.device)
    for b in range(d3.shape[0]):
        flat = d3[b].flatten()
        order = torch.argsort(flat, descending=not is_ascend, stable=True).tolist()
        used_r, used_c, n = set(), set(), 0
        M = d3.shape[2]
        for f in order:
            s = float(flat[f])
            if (s > threshold) if is_ascend else (s < threshold):
                break
            r, c = divmod(f, M)
            if r in used_r or c in used_c:
                continue
            rows[b, r] = c; cols[b, c] = r
            used_r.add(r); used_c.add(c); n += 1
            if 0 < topk <= n:
                break
    return _W(rows.reshape(shape[:-1])), _W(cols.reshape(tuple(shape[:-2]) + (shape[-1],)))


# ------------------------------------------------------------------------------------------------ SSD
def MultiBoxPrior(data, sizes=(1.0,), ratios=(1.0,), clip=False, steps=(-1.0, -1.0), offsets=(0.5, 0.5)):
    """Anchor boxes for every pixel of a feature map (``multibox_prior.cc:30-75``): per pixel ``len(sizes) + len(ratios) - 1`` boxes —
    all sizes at ``ratios[0]``, then ``sizes[0]`` at the remaining ratios.  Output ``[1, H*W*A, 4]`` (corner, normalised)."""
    x = _t(data)
    H, W = int(x.shape[-2]), int(x.shape[-1])
    sy = steps[0] if steps[0] > 0 else 1.0 / H
    sx = steps[1] if steps[1] > 0 else 1.0 / W
    cy = (torch.arange(H, dtype=torch.float32, device=x.device) + offsets[0]) * sy
    cx = (torch.arange(W, dtype=torch.float32, device=x.device) + offsets[1]) * sx
    wh = []
    r0 = _math.sqrt(ratios[0])
    for s in sizes:
        wh.append((s * H / W * r0 / 2, s / r0 / 2))
    for r in ratios[1:]:
        rr = _math.sqrt(r)
        wh.append((sizes[0] * H / W * rr / 2, sizes[0] / rr / 2))
    wh = torch.tensor(wh, dtype=torch.float32, device=x.device)                  # [A, 2] half extents
    cyx = torch.stack(torch.meshgrid(cy, cx, indexing="ij"), -1).reshape(-1, 1, 2)  # [HW, 1, (y, x)]
    out = torch.cat([cyx[..., 1:2] - wh[None, :, 0:1], cyx[..., 0:1] - wh[None, :, 1:2],
                     cyx[..., 1:2] + wh[None, :, 0:1], cyx[..., 0:1] + wh[None, :, 1:2]], -1).reshape(1, -1, 4)
    return _W(out.clamp(0, 1) if clip else out)


def _encode(anchors, gt, variances):
    a, g = _center(anchors), _center(gt)
    return torch.stack([(g[:, 0] - a[:, 0]) / a[:, 2] / variances[0], (g[:, 1] - a[:, 1]) / a[:, 3] / variances[1],
                        torch.log((g[:, 2] / a[:, 2]).clamp_min(1e-12)) / variances[2], torch.log((g[:, 3] / a[:, 3]).clamp_min(1e-12)) / variances[3]], -1)


def MultiBoxTarget(anchor, label, cls_pred, overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=-1.0,
                   negative_mining_thresh=0.5, minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2)):
    """Training targets of SSD (``multibox_target.cc:70-280``).  ``anchor [1, N, 4]``, ``label [B, M, 5]`` rows ``(cls, x1, y1, x2, y2)`` padded
    with -1, ``cls_pred [B, C, N]``.  Matching: every ground truth takes its best anchor (bipartite), then anchors with IoU >
    ``overlap_threshold`` take their best ground truth.  Returns ``[loc_target [B, 4N], loc_mask [B, 4N], cls_target [B, N]]`` with class 0 =
    background; with hard-negative mining un-mined negatives get ``ignore_label``."""
    A = _t(anchor).float().reshape(-1, 4)
    L = _t(label).float()
    P = _t(cls_pred).float()
    B, N = L.shape[0], A.shape[0]
    loc_t = torch.zeros(B, N, 4, device=A.device); loc_m = torch.zeros(B, N, 4, device=A.device)
    cls_t = torch.zeros(B, N, device=A.device)
    for b in range(B):
        gt = L[b][L[b, :, 0] >= 0]
        if gt.shape[0] == 0:
            continue
        iou = _iou(A, gt[:, 1:5])                                                  # [N, M]
        match = torch.full((N,), -1, dtype=torch.long, device=A.device)
        work = iou.clone()
        for _ in range(gt.shape[0]):                                              # bipartite stage: best remaining pair
            f = int(torch.argmax(work)); r, c = divmod(f, gt.shape[0])
            if float(work[r, c]) <= 1e-6:
                break
            match[r] = c
            work[r, :] = -1; work[:, c] = -1
        best, arg = iou.max(1)
        thr = (match < 0) & (best > overlap_threshold)
        match[thr] = arg[thr]
        pos = match >= 0
        cls_t[b, pos] = gt[match[pos], 0] + 1
        loc_t[b, pos] = _encode(A[pos], gt[match[pos], 1:5], variances)
        loc_m[b, pos] = 1
        if negative_mining_ratio > 0:
            neg = (~pos) & (best < negative_mining_thresh)
            n_neg = min(int(neg.sum()), max(int(negative_mining_ratio * int(pos.sum())), int(minimum_negative_samples)))
            bg_prob = torch.softmax(P[b], 0)[0]                                   # low background prob = hard negative
            cand = torch.nonzero(neg).flatten()
            hard = cand[torch.argsort(bg_prob[cand], stable=True)[:n_neg]]
            ign = ~pos
            ign[hard] = False
            cls_t[b, ign] = ignore_label
    return [_W(loc_t.reshape(B, -1)), _W(loc_m.reshape(B, -1)), _W(cls_t)]


def MultiBoxDetection(cls_prob, loc_pred, anchor, clip=True, threshold=0.01, background_id=0, nms_threshold=0.5, force_suppress=False,
                      variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1):
    """Decode SSD predictions (``multibox_detection.cc:45-190``): ``cls_prob [B, C, N]``, ``loc_pred [B, 4N]``, ``anchor [1, N, 4]`` →
    ``[B, N, 6]`` rows ``(class id (background removed), score, x1, y1, x2, y2)`` sorted by score, suppressed rows have id -1."""
    P = _t(cls_prob).float(); Lp = _t(loc_pred).float(); A = _center(_t(anchor).float().reshape(-1, 4))
    B, C, N = P.shape
    out = torch.full((B, N, 6), -1.0, device=P.device)
    fg = [c for c in range(C) if c != background_id]
    for b in range(B):
        score, cid = P[b][fg].max(0)
        d = Lp[b].reshape(N, 4)
        cx = d[:, 0] * variances[0] * A[:, 2] + A[:, 0]; cy = d[:, 1] * variances[1] * A[:, 3] + A[:, 1]
        w = torch.exp(d[:, 2] * variances[2]) * A[:, 2] / 2; h = torch.exp(d[:, 3] * variances[3]) * A[:, 3] / 2
        boxes = torch.stack([cx - w, cy - h, cx + w, cy + h], -1)
        if clip:
            boxes = boxes.clamp(0, 1)
        valid = torch.nonzero(score > threshold).flatten()
        order = valid[torch.argsort(score[valid], descending=True, stable=True)]
        if nms_topk > 0:
            order = order[:nms_topk]
        rows = torch.cat([cid[order].float()[:, None], score[order][:, None], boxes[order]], 1)
        if 0 < nms_threshold <= 1 and rows.shape[0]:
            keep = _nms_keep(rows[:, 2:6], rows[:, 1], rows[:, 0], nms_threshold, force_suppress)
            rows[~keep, 0] = -1
        out[b, :rows.shape[0]] = rows
    return _W(out)


# ------------------------------------------------------------------------------------------------ regions
def ROIAlign(data, rois, pooled_size, spatial_scale, sample_ratio=-1, position_sensitive=False):
    """RoIAlign (``roi_align.cc``): ``rois [R, 5]`` = (batch index, x1, y1, x2, y2); bilinear sampling, no half-pixel shift."""
    from torchvision.ops import ps_roi_align, roi_align
    ps = (pooled_size, pooled_size) if isinstance(pooled_size, int) else tuple(pooled_size)
    fn = ps_roi_align if position_sensitive else roi_align
    kw = {} if position_sensitive else {"aligned": False}
    return _W(fn(_t(data).float(), _t(rois).float(), ps, spatial_scale, sample_ratio if sample_ratio > 0 else 0, **kw))


def ROIPooling(data, rois, pooled_size, spatial_scale):
    """Max RoI pooling (``roi_pooling.cc``)."""
    from torchvision.ops import roi_pool
    ps = (pooled_size, pooled_size) if isinstance(pooled_size, int) else tuple(pooled_size)
    return _W(roi_pool(_t(data).float(), _t(rois).float(), ps, spatial_scale))


def PSROIPooling(data, rois, spatial_scale, output_dim, pooled_size, group_size=0):
    """Position-sensitive RoI pooling (``psroi_pooling.cc``)."""
    from torchvision.ops import ps_roi_pool
    out = ps_roi_pool(_t(data).float(), _t(rois).float(), (pooled_size, pooled_size), spatial_scale)
    assert out.shape[1] == output_dim, "channels must equal output_dim * pooled_size^2"
    return _W(out)


def DeformableConvolution(data, offset, weight, bias=None, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0), num_filter=None,
                          num_group=1, num_deformable_group=1, no_bias=False):
    """Deformable convolution v1 (``deformable_convolution.cc``)."""
    from torchvision.ops import deform_conv2d
    return _W(deform_conv2d(_t(data), _t(offset), _t(weight), None if (no_bias or bias is None) else _t(bias), stride=tuple(stride),
                            padding=tuple(pad), dilation=tuple(dilate)))


def _base_anchors(stride, scales, ratios, device):
    base = torch.tensor([0, 0, stride - 1, stride - 1], dtype=torch.float32, device=device)
    w = base[2] - base[0] + 1; h = base[3] - base[1] + 1
    cx = base[0] + 0.5 * (w - 1); cy = base[1] + 0.5 * (h - 1)
    out = []
    for r in ratios:
        ws = torch.round(torch.sqrt(w * h / r)); hs = torch.round(ws * r)
        for s in scales:
            W_, H_ = ws * s, hs * s
            out.append(torch.stack([cx - 0.5 * (W_ - 1), cy - 0.5 * (H_ - 1), cx + 0.5 * (W_ - 1), cy + 0.5 * (H_ - 1)]))
    return torch.stack(out)


def MultiProposal(cls_prob, bbox_pred, im_info, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7, rpn_min_size=16,
                  scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), feature_stride=16, output_score=False, iou_loss=False):
    """RPN proposals (``proposal.cc:270-420``, ``multi_proposal.cc``): anchors → apply deltas → clip → min-size filter → top-N → NMS → top-N,
    padded by repeating kept boxes.  Output ``[B*post_n, 5]`` rows ``(batch index, x1, y1, x2, y2)`` (+ scores ``[B*post_n, 1]``)."""
    from torchvision.ops import nms
    P, D, info = _t(cls_prob).float(), _t(bbox_pred).float(), _t(im_info).float()
    B, _, H, W = P.shape
    base = _base_anchors(feature_stride, scales, ratios, P.device)               # [A, 4]
    A = base.shape[0]
    sx = torch.arange(W, device=P.device) * feature_stride; sy = torch.arange(H, device=P.device) * feature_stride
    shift = torch.stack(torch.meshgrid(sy, sx, indexing="ij"), -1)               # [H, W, (y, x)]
    shifts = torch.stack([shift[..., 1], shift[..., 0], shift[..., 1], shift[..., 0]], -1).float()
    anchors = (shifts[:, :, None, :] + base[None, None]).reshape(-1, 4)          # (h, w, a) order
    rois, scores_out = [], []
    for b in range(B):
        score = P[b, A:].permute(1, 2, 0).reshape(-1)
        d = D[b].reshape(A, 4, H, W).permute(2, 3, 0, 1).reshape(-1, 4)
        if iou_loss:
            boxes = anchors + d
        else:
            w = anchors[:, 2] - anchors[:, 0] + 1; h = anchors[:, 3] - anchors[:, 1] + 1
            cx = anchors[:, 0] + 0.5 * (w - 1); cy = anchors[:, 1] + 0.5 * (h - 1)
            pcx = d[:, 0] * w + cx; pcy = d[:, 1] * h + cy; pw = torch.exp(d[:, 2]) * w; ph = torch.exp(d[:, 3]) * h
            boxes = torch.stack([pcx - 0.5 * (pw - 1), pcy - 0.5 * (ph - 1), pcx + 0.5 * (pw - 1), pcy + 0.5 * (ph - 1)], -1)
        ih, iw, sc = float(info[b, 0]), float(info[b, 1]), float(info[b, 2])
        boxes = torch.stack([boxes[:, 0].clamp(0, iw - 1), boxes[:, 1].clamp(0, ih - 1), boxes[:, 2].clamp(0, iw - 1), boxes[:, 3].clamp(0, ih - 1)], -1)
        ms = rpn_min_size * sc
        ok = ((boxes[:, 2] - boxes[:, 0] + 1) >= ms) & ((boxes[:, 3] - boxes[:, 1] + 1) >= ms)
        score = torch.where(ok, score, torch.full_like(score, -1.0))
        order = torch.argsort(score, descending=True, stable=True)[:rpn_pre_nms_top_n if rpn_pre_nms_top_n > 0 else None]
        bx, sc_ = boxes[order], score[order]
        keep = nms(bx, sc_, threshold)[:rpn_post_nms_top_n]
        if keep.numel() < rpn_post_nms_top_n:                                    # pad by cycling through the kept ones
            keep = keep[torch.arange(rpn_post_nms_top_n, device=keep.device) % max(1, keep.numel())]
        rois.append(torch.cat([torch.full((rpn_post_nms_top_n, 1), float(b), device=P.device), bx[keep]], 1))
        scores_out.append(sc_[keep][:, None])
    r = _W(torch.cat(rois, 0))
    return (r, _W(torch.cat(scores_out, 0))) if output_score else r


Proposal = MultiProposal


# ------------------------------------------------------------------------------------------------ signal / sketch
def fft(data, compute_size=128):
    """Real → interleaved complex FFT over the last axis: ``[..., d]`` → ``[..., 2d]`` (re, im pairs) (``contrib/fft-inl.h``)."""
    c = torch.fft.fft(_t(data).float(), dim=-1)
    return _W(torch.view_as_real(c).reshape(*c.shape[:-1], -1))


def ifft(data, compute_size=128):
    """Interleaved complex ``[..., 2d]`` → real part ``[..., d]`` of the UNNORMALISED inverse transform, like cuFFT (``contrib/ifft-inl.h``):
    ``ifft(fft(x)) == d * x``."""
    x = _t(data).float()
    c = torch.view_as_complex(x.reshape(*x.shape[:-1], -1, 2).contiguous())
    return _W(torch.fft.ifft(c, dim=-1).real * c.shape[-1])


def count_sketch(data, h, s, out_dim, processing_batch_size=32):
    """Count sketch projection (``contrib/count_sketch-inl.h``): ``out[n, h[i]] += s[i] * data[n, i]``."""
    x = _t(data).float(); hh = _t(h).long().reshape(-1); ss = _t(s).float().reshape(-1)
    out = torch.zeros(x.shape[0], int(out_dim), dtype=x.dtype, device=x.device)
    return _W(out.index_add(1, hh, x * ss[None, :]))


# ------------------------------------------------------------------------------------------------ quantisation
def _qrange(out_type):
    return {"uint8": (0.0, 255.0, torch.uint8), "int8": (-127.0, 127.0, torch.int8), "int32": (-2147483647.0, 2147483647.0, torch.int32)}[out_type]


def quantize(data, min_range, max_range, out_type="uint8"):
    """Affine (uint8) / symmetric (int8) quantisation (``quantization/quantize-inl.h``).  Returns ``(q, min, max)``."""
    x = _t(data).float(); lo = _t(min_range).float().reshape(()); hi = _t(max_range).float().reshape(())
    qlo, qhi, dt = _qrange(out_type)
    if out_type == "uint8":
        scale = (qhi - qlo) / (hi - lo).clamp_min(1e-30)
        q = torch.round((x - lo) * scale).clamp(qlo, qhi).to(dt)
        return _W(q), _W(lo.reshape(1)), _W(hi.reshape(1))
    r = torch.maximum(lo.abs(), hi.abs())
    q = (torch.sign(x) * torch.floor(x.abs() * (qhi / r.clamp_min(1e-30)) + 0.5)).clamp(qlo, qhi).to(dt)
    return _W(q), _W((-r).reshape(1)), _W(r.reshape(1))


def dequantize(data, min_range, max_range, out_type="float32"):
    q = _t(data); lo = _t(min_range).float().reshape(()); hi = _t(max_range).float().reshape(())
    if q.dtype == torch.uint8:
        return _W(q.float() * ((hi - lo) / 255.0) + lo)
    qmax = 127.0 if q.dtype == torch.int8 else 2147483647.0
    return _W(q.float() * (torch.maximum(lo.abs(), hi.abs()) / qmax))


def requantize(data, min_range, max_range, min_calib_range=None, max_calib_range=None):
    """int32 accumulators → int8 (``quantization/requantize-inl.h``): range from calibration or from the data itself."""
    real = _t(dequantize(data, min_range, max_range))
    if min_calib_range is not None and max_calib_range is not None:
        r = max(abs(float(min_calib_range)), abs(float(max_calib_range)))
    else:
        r = float(real.abs().max())
    return quantize(_W(real), _W(torch.tensor(-r)), _W(torch.tensor(r)), "int8")


# ------------------------------------------------------------------------------------------------ misc
def boolean_mask(data, index, axis=0):
    x = _t(data); m = _t(index) != 0
    return _W(x.movedim(axis, 0)[m].movedim(0, axis) if axis else x[m])


def index_array(data, axes=None):
    x = _t(data)
    grids = torch.meshgrid(*[torch.arange(s, device=x.device) for s in x.shape], indexing="ij")
    out = torch.stack(grids, -1)
    return _W(out if axes is None else out[..., list(axes)])


def getnnz(data, axis=None):
    x = data.tostype("default")._t if hasattr(data, "tostype") and getattr(data, "stype", "default") != "default" else _t(data)
    return _W((x != 0).sum() .reshape(1) if axis is None else (x != 0).sum(axis))


class _GradMul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scalar):
        ctx.s = scalar
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.s, None


def gradientmultiplier(data, scalar=1.0):
    """Identity forward, gradient × ``scalar`` backward (gradient reversal with a negative scalar)."""
    return _W(_GradMul.apply(_t(data), float(scalar)))


def isnan(data): return _W(torch.isnan(_t(data)).to(_t(data).dtype))
def isinf(data): return _W(torch.isinf(_t(data)).to(_t(data).dtype))
def isfinite(data): return _W(torch.isfinite(_t(data)).to(_t(data).dtype))


def SyncBatchNorm(data, gamma, beta, moving_mean, moving_var, eps=1e-3, momentum=0.9, fix_gamma=False, use_global_stats=False,
                  ndev=1, key=""):
    """Functional cross-rank BatchNorm: statistics are all-reduced over the default process group when one exists
    (``contrib/sync_batch_norm-inl.h:79-455`` uses host-side shared accumulators + a barrier)."""
    import torch.distributed as dist
    x = _t(data)
    red = [0] + list(range(2, x.dim()))
    if use_global_stats:
        mean, var = _t(moving_mean), _t(moving_var)
    else:
        n = torch.tensor([float(x.numel() // x.shape[1])], device=x.device)
        s1 = x.sum(red); s2 = (x * x).sum(red)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = torch.cat([s1.detach(), s2.detach(), n])
            dist.all_reduce(packed)
            C = s1.numel()
            # gradients flow through the local terms; the remote part is a constant of this rank's graph
            s1 = s1 + (packed[:C] - s1.detach()); s2 = s2 + (packed[C:2 * C] - s2.detach()); n = packed[2 * C:]
        mean = s1 / n; var = (s2 / n - mean * mean).clamp_min(0)
        with torch.no_grad():
            _t(moving_mean).mul_(momentum).add_(mean.detach() * (1 - momentum))
            _t(moving_var).mul_(momentum).add_(var.detach() * (1 - momentum))
    shp = [1, -1] + [1] * (x.dim() - 2)
    g = torch.ones_like(_t(gamma)) if fix_gamma else _t(gamma)
    return _W((x - mean.view(shp)) * torch.rsqrt(var.view(shp) + eps) * g.view(shp) + _t(beta).view(shp))


# ------------------------------------------------------------------------------------------------ control flow
def _flat(x):
    return [x] if isinstance(x, NDArray) else list(x)


def foreach(body, data, init_states):
    """``body(slice_t, states) -> (out_t, new_states)`` scanned over axis 0 of ``data`` (``ndarray/contrib.py:100-220``).  Imperative: the
    loop simply runs; outputs are stacked."""
    single = isinstance(data, NDArray)
    seqs = _flat(data)
    states = init_states
    outs = []
    for t in range(seqs[0].shape[0]):
        xt = seqs[0][t] if single else [s[t] for s in seqs]
        o, states = body(xt, states)
        outs.append(o)
    if not outs:
        return [], states
    if isinstance(outs[0], NDArray):
        return _W(torch.stack([o._t for o in outs])), states
    return [_W(torch.stack([o[i]._t for o in outs])) for i in range(len(outs[0]))], states


def while_loop(cond, func, loop_vars, max_iterations=None):
    """``while cond(*vars): out, vars = func(*vars)`` with at most ``max_iterations`` steps; outputs are stacked and zero-padded to
    ``max_iterations`` rows like the reference (``ndarray/contrib.py:232-390``)."""
    if max_iterations is None:
        raise ValueError("max_iterations should be specified")
    single = isinstance(loop_vars, NDArray)
    vars_ = _flat(loop_vars)
    outs, steps = [], 0
    while steps < max_iterations and bool(_t(cond(*vars_)).reshape(-1)[0] != 0):
        o, nv = func(*vars_)
        vars_ = _flat(nv)
        if o is not None:
            outs.append(_flat(o))
        steps += 1
    stacked = []
    if outs:
        for i in range(len(outs[0])):
            s = torch.stack([o[i]._t for o in outs])
            if steps < max_iterations:
                s = torch.cat([s, torch.zeros((max_iterations - steps,) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)])
            stacked.append(_W(s))
    return stacked, (vars_[0] if single else vars_)


def cond(pred, then_func, else_func):
    """``then_func()`` if ``pred`` is non-zero else ``else_func()`` (``ndarray/contrib.py:400-460``)."""
    return then_func() if bool(_t(pred).reshape(-1)[0] != 0) else else_func()



# ------------------------------------------------------------------------------------------------ remaining contrib registrations
def SparseEmbedding(data, weight, input_dim=None, output_dim=None, dtype="float32", deterministic=False):
    """Embedding lookup whose weight gradient is row-sparse in the reference; the lookup itself is identical."""
    return _W(TF.embedding(_t(data).long(), _t(weight)))


def group_adagrad_update(weight, grad, history, lr, rescale_grad=1.0, clip_gradient=-1.0, epsilon=1e-5, out=None):
    """AdaGrad with ONE accumulator per row (``contrib/optimizer_op-inl.h``): ``history += mean(g^2, axis=1)``;
    ``w -= lr * g / sqrt(history + eps)``.  In place on ``weight`` / ``history``."""
    with torch.no_grad():
        g = _t(grad) * rescale_grad
        if clip_gradient is not None and clip_gradient >= 0:
            g = g.clamp(-clip_gradient, clip_gradient)
        h = _t(history)
        h.add_((g * g).mean(dim=tuple(range(1, g.dim())), keepdim=True).reshape(h.shape))
        new = _t(weight) - lr * g / torch.sqrt(h.reshape([-1] + [1] * (g.dim() - 1)) + epsilon)
        tgt = weight if out is None else out
        _t(tgt).copy_(new)
        return tgt


def _q2f(data, lo, hi):
    return _t(dequantize(data, lo, hi))


def _requant_out(real, out_type="int8"):
    r = float(real.abs().max()) if real.numel() else 0.0
    return quantize(_W(real), _W(torch.tensor(-r)), _W(torch.tensor(r)), out_type)


def quantized_fully_connected(data, weight, bias, min_data, max_data, min_weight, max_weight, min_bias=None, max_bias=None, num_hidden=None,
                              no_bias=False, flatten=True):
    """int8 FullyConnected, simulated: operands are dequantised, multiplied in fp32 and the result returned as int32-range accumulators
    ``(out, min_out, max_out)`` like ``quantized_fully_connected.cc`` (the float value of an accumulator unit is
    ``scale_data * scale_weight``)."""
    x, w = _t(data).float(), _t(weight).float()
    sd = float(torch.maximum(_t(min_data).abs().max(), _t(max_data).abs().max())) / 127.0
    sw = float(torch.maximum(_t(min_weight).abs().max(), _t(max_weight).abs().max())) / 127.0
    acc = TF.linear(x.flatten(1) if flatten else x, w)
    if not no_bias and bias is not None:
        sb = float(torch.maximum(_t(min_bias).abs().max(), _t(max_bias).abs().max())) / 127.0
        acc = acc + torch.round(_t(bias).float() * sb / (sd * sw))
    rng = 2147483647.0 * sd * sw
    return _W(acc.to(torch.int32)), _W(torch.tensor([-rng])), _W(torch.tensor([rng]))


def quantized_conv(data, weight, bias, min_data, max_data, min_weight, max_weight, min_bias=None, max_bias=None, kernel=None, stride=(1, 1),
                   pad=(0, 0), dilate=(1, 1), num_filter=None, num_group=1, no_bias=True, layout=None):
    x, w = _t(data).float(), _t(weight).float()
    sd = float(torch.maximum(_t(min_data).abs().max(), _t(max_data).abs().max())) / 127.0
    sw = float(torch.maximum(_t(min_weight).abs().max(), _t(max_weight).abs().max())) / 127.0
    acc = TF.conv2d(x, w, None, tuple(stride), tuple(pad), tuple(dilate), num_group)
    if not no_bias and bias is not None:
        sb = float(torch.maximum(_t(min_bias).abs().max(), _t(max_bias).abs().max())) / 127.0
        acc = acc + torch.round(_t(bias).float() * sb / (sd * sw)).view(1, -1, 1, 1)
    rng = 2147483647.0 * sd * sw
    return _W(acc.to(torch.int32)), _W(torch.tensor([-rng])), _W(torch.tensor([rng]))


def quantized_pooling(data, min_data, max_data, kernel=(2, 2), pool_type="max", stride=None, pad=(0, 0), global_pool=False, **kw):
    """Pooling directly on the int8 codes (max / avg commute with the affine map); ranges pass through."""
    x = _t(data)
    k = tuple(x.shape[2:]) if global_pool else tuple(kernel)
    st = tuple(stride) if stride else k
    y = TF.max_pool2d(x.float(), k, st, tuple(pad)) if pool_type == "max" else torch.round(TF.avg_pool2d(x.float(), k, st, tuple(pad)))
    return _W(y.to(x.dtype)), min_data, max_data


def quantized_flatten(data, min_data, max_data):
    return _W(_t(data).flatten(1)), min_data, max_data


def quantized_concat(*args, dim=1, num_args=None):
    """``quantized_concat(d0, d1, …, min0, max0, min1, max1, …)``: inputs are re-scaled to the widest range, then concatenated."""
    n = num_args or len(args) // 3
    datas, ranges = args[:n], args[n:]
    rs = [float(torch.maximum(_t(ranges[2 * i]).abs().max(), _t(ranges[2 * i + 1]).abs().max())) for i in range(n)]
    top = max(rs)
    parts = [torch.round(_t(d).float() * (r / top)).clamp(-127, 127).to(_t(d).dtype) for d, r in zip(datas, rs)]
    return _W(torch.cat(parts, dim=dim)), _W(torch.tensor([-top])), _W(torch.tensor([top]))


# ---- DGL graph helpers on CSR adjacency matrices whose stored values are edge ids (contrib/dgl_graph.cc)
def edge_id(data, u, v):
    """Edge id ``data[u[i], v[i]]`` or -1 when the edge does not exist."""
    ptr, idx, val = data.indptr._t.long(), data.indices._t.long(), data.data._t
    out = torch.full((_t(u).numel(),), -1.0, dtype=val.dtype if val.is_floating_point() else torch.float32)
    for i, (a, b) in enumerate(zip(_t(u).long().tolist(), _t(v).long().tolist())):
        cols = idx[ptr[a]:ptr[a + 1]]
        hit = torch.nonzero(cols == b)
        if hit.numel():
            out[i] = val[ptr[a] + hit[0, 0]]
    return _W(out)


def dgl_adjacency(data):
    """Same sparsity pattern with every stored value replaced by 1.0."""
    from .sparse import CSRNDArray
    return CSRNDArray(_W(torch.ones(data.data._t.shape, dtype=torch.float32)), data.indices.copy(), data.indptr.copy(), data.shape)


def dgl_subgraph(graph, *vertex_sets, return_mapping=False, num_args=None):
    """Vertex-induced subgraphs: for every id list ``v`` the CSR of ``graph[v][:, v]`` with edges renumbered from 0 (and, with
    ``return_mapping``, a second CSR of the same pattern holding the ORIGINAL edge ids)."""
    from .sparse import CSRNDArray
    ptr, idx, val = graph.indptr._t.long(), graph.indices._t.long(), graph.data._t
    subs, maps = [], []
    for vs in vertex_sets:
        ids = _t(vs).long().tolist()
        remap = {v: i for i, v in enumerate(ids)}
        nptr, nidx, orig = [0], [], []
        for v in ids:
            for j in range(int(ptr[v]), int(ptr[v + 1])):
                c = int(idx[j])
                if c in remap:
                    nidx.append(remap[c]); orig.append(float(val[j]))
            nptr.append(len(nidx))
        shape = (len(ids), len(ids))
        mk = lambda d: CSRNDArray(_W(torch.tensor(d, dtype=torch.float32)), _W(torch.tensor(nidx, dtype=torch.int64)),  # noqa: E731
                                  _W(torch.tensor(nptr, dtype=torch.int64)), shape)
        subs.append(mk(list(range(len(nidx))))); maps.append(mk(orig))
    res = subs + (maps if return_mapping else [])
    return res[0] if len(res) == 1 else res


def DeformablePSROIPooling(data, rois, trans=None, spatial_scale=1.0, output_dim=None, group_size=None, pooled_size=None, part_size=0,
                           sample_per_part=1, trans_std=0.0, no_trans=False):
    """Deformable position-sensitive RoI pooling (``contrib/deformable_psroi_pooling.cu``): every output bin ``(c, ph, pw)`` averages
    ``sample_per_part²`` bilinear samples of the position-sensitive channel ``(c·G + gh)·G + gw``, the bin being shifted by the learned,
    RoI-size-normalised offset ``trans[r, 2·class (+1), part_h, part_w] · trans_std``.  ``data [N, output_dim·G², H, W]``, ``rois [R, 5]``."""
    x, r = _t(data).float(), _t(rois).float()
    N, C, H, W = x.shape
    P, G, S = int(pooled_size), int(group_size), int(sample_per_part)
    part = int(part_size) if part_size else P
    R = r.shape[0]
    out = x.new_zeros((R, int(output_dim), P, P))
    tr = None if (no_trans or trans is None) else _t(trans).float()
    ncls = 1 if tr is None else tr.shape[1] // 2
    ch_per_cls = int(output_dim) // ncls
    ph = torch.arange(P, device=x.device).view(1, P, 1, 1, 1).float(); pw = torch.arange(P, device=x.device).view(1, 1, P, 1, 1).float()
    ih = torch.arange(S, device=x.device).view(1, 1, 1, S, 1).float(); iw = torch.arange(S, device=x.device).view(1, 1, 1, 1, S).float()
    ctop = torch.arange(int(output_dim), device=x.device).view(-1, 1, 1, 1, 1)
    gh = torch.clamp(torch.floor(ph * G / P), 0, G - 1).long(); gw = torch.clamp(torch.floor(pw * G / P), 0, G - 1).long()
    chan = ((ctop * G + gh) * G + gw).expand(-1, P, P, S, S)                              # [D, P, P, S, S]
    part_h = torch.floor(ph / P * part).long().clamp(0, part - 1); part_w = torch.floor(pw / P * part).long().clamp(0, part - 1)
    cls = (ctop // ch_per_cls).clamp(0, ncls - 1)
    for i in range(R):
        b = int(r[i, 0])
        sw = torch.round(r[i, 1]) * spatial_scale - 0.5; sh = torch.round(r[i, 2]) * spatial_scale - 0.5
        ew = (torch.round(r[i, 3]) + 1.0) * spatial_scale - 0.5; eh = (torch.round(r[i, 4]) + 1.0) * spatial_scale - 0.5
        rw = torch.clamp(ew - sw, min=0.1); rh = torch.clamp(eh - sh, min=0.1)
        bw, bh = rw / P, rh / P
        if tr is None:
            tx = ty = x.new_zeros(())
        else:
            tx = tr[i][cls * 2, part_h, part_w] * trans_std                               # broadcast to [D, P, P, 1, 1]
            ty = tr[i][cls * 2 + 1, part_h, part_w] * trans_std
        ws = pw * bw + sw + tx * rw + iw * (bw / S)
        hs = ph * bh + sh + ty * rh + ih * (bh / S)
        ws, hs = ws.expand(int(output_dim), P, P, S, S), hs.expand(int(output_dim), P, P, S, S)
        ok = (ws >= -0.5) & (ws <= W - 0.5) & (hs >= -0.5) & (hs <= H - 0.5)
        wc, hc = ws.clamp(0, W - 1), hs.clamp(0, H - 1)
        x0, y0 = torch.floor(wc).long(), torch.floor(hc).long()
        x1, y1 = (x0 + 1).clamp(max=W - 1), (y0 + 1).clamp(max=H - 1)
        fx, fy = wc - x0, hc - y0
        img = x[b]
        v = (img[chan, y0, x0] * (1 - fx) * (1 - fy) + img[chan, y0, x1] * fx * (1 - fy) + img[chan, y1, x0] * (1 - fx) * fy + img[chan, y1, x1] * fx * fy)
        cnt = ok.sum(dim=(-1, -2)).clamp_min(1)
        out[i] = (v * ok).sum(dim=(-1, -2)) / cnt
    return _W(out)

__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "TF", "NDArray", "annotations")]
