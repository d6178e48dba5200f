"""Loss blocks.  Parity: ``python/mxnet/gluon/loss.py`` — ``_apply_weighting``, ``_reshape_like``, L2Loss, L1Loss,
SigmoidBinaryCrossEntropyLoss, SoftmaxCrossEntropyLoss (:304-318: ``-pick(log_softmax(pred), label)`` then mean
over non-batch axes), KLDivLoss, HuberLoss, HingeLoss, SquaredHingeLoss, LogisticLoss, CTCLoss (:437-520), TripletLoss (:640-690), PoissonNLLLoss (:700-760),
CosineEmbeddingLoss (:770-830)."""
from __future__ import annotations

import torch

from ..ndarray import NDArray
from ..ops import functional as OF
from .block import HybridBlock

__all__ = ["Loss", "L2Loss", "L1Loss", "SigmoidBinaryCrossEntropyLoss", "SigmoidBCELoss", "SoftmaxCrossEntropyLoss",
           "SoftmaxCELoss", "KLDivLoss", "HuberLoss", "HingeLoss", "SquaredHingeLoss", "LogisticLoss", "CTCLoss", "TripletLoss",
           "PoissonNLLLoss", "CosineEmbeddingLoss"]


def _w(loss, weight, sample_weight):
    if sample_weight is not None:
        loss = loss * sample_weight._t
    if weight is not None:
        loss = loss * weight
    return loss


def _mean(loss, batch_axis):
    dims = [d for d in range(loss.dim()) if d != batch_axis]
    return loss.mean(dim=dims) if dims else loss


class Loss(HybridBlock):
    def __init__(self, weight, batch_axis, **kwargs):
        super().__init__(**kwargs)
        self._weight, self._batch_axis = weight, batch_axis

    def forward(self, x, *args):
        return self.hybrid_forward(None, x, *args)


class L2Loss(Loss):
    def __init__(self, weight=1.0, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs)

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        l = (pred._t - label._t.reshape(pred._t.shape)).square()
        return NDArray(_mean(_w(l, self._weight / 2, sample_weight), self._batch_axis))


class L1Loss(Loss):
    def __init__(self, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs)

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        l = (pred._t - label._t.reshape(pred._t.shape)).abs()
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


class SigmoidBinaryCrossEntropyLoss(Loss):
    def __init__(self, from_sigmoid=False, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._from_sigmoid = from_sigmoid

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        p, y = pred._t, label._t.reshape(pred._t.shape)
        if not self._from_sigmoid:
            l = torch.relu(p) - p * y + torch.nn.functional.softplus(-p.abs())
        else:
            l = -(torch.log(p + 1e-12) * y + torch.log(1. - p + 1e-12) * (1. - y))
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


SigmoidBCELoss = SigmoidBinaryCrossEntropyLoss


class SoftmaxCrossEntropyLoss(Loss):
    def __init__(self, axis=-1, sparse_label=True, from_logits=False, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs)
        self._axis, self._sparse_label, self._from_logits = axis, sparse_label, from_logits

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        p, y = pred._t, label._t
        if self._from_logits:
            if self._sparse_label:
                l = -torch.gather(p, self._axis, y.long().unsqueeze(self._axis)).squeeze(self._axis)
            else:
                l = -(p * y).sum(dim=self._axis)
        else:
            l = OF.softmax_cross_entropy(p, y, self._sparse_label, self._axis)
        l = _w(l, self._weight, sample_weight)
        return NDArray(_mean(l, self._batch_axis) if l.dim() > 1 else l)


SoftmaxCELoss = SoftmaxCrossEntropyLoss


class KLDivLoss(Loss):
    def __init__(self, from_logits=True, axis=-1, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._from_logits, self._axis = from_logits, axis

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        p = pred._t if self._from_logits else torch.log_softmax(pred._t, self._axis)
        l = label._t * (torch.log(label._t + 1e-12) - p)
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


class HuberLoss(Loss):
    def __init__(self, rho=1, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._rho = rho

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        l = (pred._t - label._t.reshape(pred._t.shape)).abs()
        l = torch.where(l > self._rho, l - 0.5 * self._rho, (0.5 / self._rho) * l.square())
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


class HingeLoss(Loss):
    def __init__(self, margin=1, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._margin = margin

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        l = torch.relu(self._margin - pred._t * label._t.reshape(pred._t.shape))
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


class SquaredHingeLoss(HingeLoss):
    def hybrid_forward(self, F, pred, label, sample_weight=None):
        l = torch.relu(self._margin - pred._t * label._t.reshape(pred._t.shape)).square()
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


class LogisticLoss(Loss):
    def __init__(self, weight=None, batch_axis=0, label_format="signed", **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._fmt = label_format

    def hybrid_forward(self, F, pred, label, sample_weight=None):
        y = label._t.reshape(pred._t.shape)
        if self._fmt == "signed":
            y = (y + 1.0) / 2.0
        l = torch.relu(pred._t) - pred._t * y + torch.nn.functional.softplus(-pred._t.abs())
        return NDArray(_mean(_w(l, self._weight, sample_weight), self._batch_axis))


class CTCLoss(Loss):
    """Connectionist temporal classification loss.  ``layout`` 'NTC' / 'TNC' for predictions (un-normalised activations; the blank is the
    LAST class, as in gluon), ``label_layout`` 'NT' / 'TN'; labels are padded with -1 unless ``label_lengths`` is given."""

    def __init__(self, layout="NTC", label_layout="NT", weight=None, **kwargs):
        assert layout in ("NTC", "TNC") and label_layout in ("NT", "TN")
        super().__init__(weight, label_layout.find("N"), **kwargs)
        self._layout, self._label_layout = layout, label_layout

    def hybrid_forward(self, F, pred, label, pred_lengths=None, label_lengths=None, sample_weight=None):
        p = pred._t if self._layout == "TNC" else pred._t.transpose(0, 1)
        lab = label._t if self._label_layout == "NT" else label._t.transpose(0, 1)
        T, N, C = p.shape
        il = pred_lengths._t.long() if pred_lengths is not None else torch.full((N,), T, dtype=torch.long)
        ll = label_lengths._t.long() if label_lengths is not None else (lab >= 0).sum(1)
        loss = torch.nn.functional.ctc_loss(torch.log_softmax(p.float(), -1), lab.long().clamp_min(0), il, ll, blank=C - 1, reduction="none", zero_infinity=False)
        return NDArray(_w(loss, self._weight, sample_weight))


class TripletLoss(Loss):
    """``max(sum((pred - pos)^2 - (pred - neg)^2) + margin, 0)``"""

    def __init__(self, margin=1, weight=None, batch_axis=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._margin = margin

    def hybrid_forward(self, F, pred, positive, negative, sample_weight=None):
        d = ((pred._t - positive._t.reshape(pred._t.shape)) ** 2 - (pred._t - negative._t.reshape(pred._t.shape)) ** 2)
        dims = [i for i in range(d.dim()) if i != self._batch_axis]
        loss = torch.relu(d.sum(dims) + self._margin)
        return NDArray(_w(loss, self._weight, sample_weight))


class PoissonNLLLoss(Loss):
    """Poisson negative log likelihood; ``from_logits`` (default) means ``pred`` is log-rate.  Returns the MEAN over all elements."""

    def __init__(self, weight=None, from_logits=True, batch_axis=0, compute_full=False, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._from_logits, self._full = from_logits, compute_full

    def hybrid_forward(self, F, pred, target, sample_weight=None, epsilon=1e-08):
        p, t = pred._t, target._t.reshape(pred._t.shape)
        loss = torch.exp(p) - t * p if self._from_logits else p - t * torch.log(p + epsilon)
        if self._full:
            st = t * torch.log(t.clamp_min(1e-30)) - t + 0.5 * torch.log(2 * t.clamp_min(1e-30) * 3.141592653589793)
            loss = loss + torch.where(t > 1, st, torch.zeros_like(st))
        return NDArray(_w(loss, self._weight, sample_weight).mean())


class CosineEmbeddingLoss(Loss):
    """``1 - cos(a, b)`` for label 1, ``max(0, cos(a, b) - margin)`` for label -1."""

    def __init__(self, weight=None, batch_axis=0, margin=0, **kwargs):
        super().__init__(weight, batch_axis, **kwargs); self._margin = margin

    def hybrid_forward(self, F, input1, input2, label, sample_weight=None):
        a, b = input1._t, input2._t.reshape(input1._t.shape)
        cos = (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1)).clamp_min(1e-12)
        lab = label._t.reshape(cos.shape)
        loss = torch.where(lab == 1, 1 - cos, torch.relu(cos - self._margin))
        return NDArray(_w(loss, self._weight, sample_weight))
