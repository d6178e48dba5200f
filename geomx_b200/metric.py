"""Evaluation metrics.  Parity: ``python/mxnet/metric.py`` (EvalMetric, Accuracy, TopKAccuracy, MAE, MSE, RMSE,
CrossEntropy, Loss, CompositeEvalMetric, create)."""
from __future__ import annotations

import math

import torch

__all__ = ["EvalMetric", "Accuracy", "TopKAccuracy", "MAE", "MSE", "RMSE", "CrossEntropy", "NegativeLogLikelihood", "Loss", "Torch", "Caffe",
           "CompositeEvalMetric", "F1", "MCC", "Perplexity", "PearsonCorrelation", "CustomMetric", "np", "create", "register", "alias",
           "check_label_shapes"]


def _t(x):
    return x._t.detach() if hasattr(x, "_t") else torch.as_tensor(x)


class EvalMetric:
    def __init__(self, name, output_names=None, label_names=None, **kwargs):
        self.name = str(name); self.output_names, self.label_names = output_names, label_names
        self._kwargs = kwargs
        self.reset()

    def reset(self):
        self.num_inst, self.sum_metric = 0, 0.0

    def update(self, labels, preds):
        raise NotImplementedError

    def get(self):
        return (self.name, float("nan")) if self.num_inst == 0 else (self.name, self.sum_metric / self.num_inst)

    def get_config(self):
        """JSON-able description: ``create(**config)`` rebuilds the metric (metric.py:95-107)."""
        config = dict(self._kwargs)
        config.update({"metric": self.__class__.__name__, "name": self.name, "output_names": self.output_names, "label_names": self.label_names})
        return config

    def update_dict(self, label, pred):
        """Update from ``{name: array}`` dicts, selecting ``output_names`` / ``label_names`` when they were given (metric.py:109-130)."""
        preds = [pred[n] for n in self.output_names] if self.output_names is not None else list(pred.values())
        labels = [label[n] for n in self.label_names] if self.label_names is not None else list(label.values())
        self.update(labels, preds)

    def get_name_value(self):
        name, value = self.get()
        if not isinstance(name, list):
            name, value = [name], [value]
        return list(zip(name, value))

    def __str__(self):
        return "EvalMetric: {}".format(dict(self.get_name_value()))


def _lists(labels, preds):
    if not isinstance(labels, (list, tuple)):
        labels = [labels]
    if not isinstance(preds, (list, tuple)):
        preds = [preds]
    return labels, preds


class Accuracy(EvalMetric):
    def __init__(self, axis=1, name="accuracy", **kw):
        super().__init__(name, **kw); self.axis = axis

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l), _t(p)
            if p.dim() > l.dim():
                p = p.argmax(dim=self.axis)
            self.sum_metric += float((p.reshape(-1).long() == l.reshape(-1).long().to(p.device)).sum())
            self.num_inst += l.numel()


class TopKAccuracy(EvalMetric):
    def __init__(self, top_k=1, name="top_k_accuracy", **kw):
        super().__init__(name + "_%d" % top_k, **kw); self.top_k = top_k

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l), _t(p)
            top = p.topk(self.top_k, dim=1).indices
            self.sum_metric += float((top == l.reshape(-1, 1).long().to(p.device)).any(dim=1).sum())
            self.num_inst += l.numel()


class MAE(EvalMetric):
    def __init__(self, name="mae", **kw):
        super().__init__(name, **kw)

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).float(), _t(p).float()
            self.sum_metric += float((l.reshape(p.shape).to(p.device) - p).abs().mean()); self.num_inst += 1


class MSE(EvalMetric):
    def __init__(self, name="mse", **kw):
        super().__init__(name, **kw)

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).float(), _t(p).float()
            self.sum_metric += float((l.reshape(p.shape).to(p.device) - p).square().mean()); self.num_inst += 1


class RMSE(MSE):
    def __init__(self, name="rmse", **kw):
        super().__init__(name, **kw)

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).float(), _t(p).float()
            self.sum_metric += math.sqrt(float((l.reshape(p.shape).to(p.device) - p).square().mean())); self.num_inst += 1


class CrossEntropy(EvalMetric):
    def __init__(self, eps=1e-12, name="cross-entropy", **kw):
        super().__init__(name, **kw); self.eps = eps

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).reshape(-1).long(), _t(p)
            prob = p[torch.arange(l.numel(), device=p.device), l.to(p.device)]
            self.sum_metric += float((-torch.log(prob + self.eps)).sum()); self.num_inst += l.numel()


class Loss(EvalMetric):
    def __init__(self, name="loss", **kw):
        super().__init__(name, **kw)

    def update(self, _, preds):
        if not isinstance(preds, (list, tuple)):
            preds = [preds]
        for p in preds:
            p = _t(p); self.sum_metric += float(p.sum()); self.num_inst += p.numel()


class CompositeEvalMetric(EvalMetric):
    def __init__(self, metrics=None, name="composite", **kw):
        self.metrics = [create(m) for m in (metrics or [])]
        super().__init__(name, **kw)

    def add(self, metric):
        self.metrics.append(create(metric))

    def reset(self):
        for m in getattr(self, "metrics", []):
            m.reset()

    def update(self, labels, preds):
        for m in self.metrics:
            m.update(labels, preds)

    def get(self):
        names, values = [], []
        for m in self.metrics:
            n, v = m.get(); names.append(n); values.append(v)
        return names, values


class F1(EvalMetric):
    """Binary F1 (python/mxnet/metric.py F1 :560-650): ``average='macro'`` averages the per-batch scores, ``'micro'`` pools the counts."""

    def __init__(self, name="f1", average="macro", **kw):
        self.average = average
        super().__init__(name, **kw)

    def reset(self):
        super().reset()
        self.tp = self.fp = self.fn = 0.0

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).reshape(-1).long(), _t(p)
            pred = p.argmax(dim=1) if p.dim() > 1 else (p > 0.5).long()
            l = l.to(pred.device)
            tp, fp, fn = float(((pred == 1) & (l == 1)).sum()), float(((pred == 1) & (l == 0)).sum()), float(((pred == 0) & (l == 1)).sum())
            if self.average == "macro":
                prec, rec = tp / max(tp + fp, 1e-12), tp / max(tp + fn, 1e-12)
                self.sum_metric += 2 * prec * rec / max(prec + rec, 1e-12); self.num_inst += 1
            else:
                self.tp += tp; self.fp += fp; self.fn += fn
                prec, rec = self.tp / max(self.tp + self.fp, 1e-12), self.tp / max(self.tp + self.fn, 1e-12)
                self.sum_metric, self.num_inst = 2 * prec * rec / max(prec + rec, 1e-12), 1


class Perplexity(EvalMetric):
    """exp(mean negative log-likelihood) with an optional ``ignore_label`` (metric.py Perplexity :760-840)."""

    def __init__(self, ignore_label=None, axis=-1, name="perplexity", **kw):
        self.ignore_label, self.axis = ignore_label, axis
        super().__init__(name, **kw)

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).reshape(-1).long(), _t(p)
            p = p.reshape(-1, p.shape[-1])
            prob = p[torch.arange(l.numel(), device=p.device), l.to(p.device)]
            if self.ignore_label is not None:
                keep = l.to(p.device) != self.ignore_label
                prob = prob[keep]
            self.sum_metric += float((-torch.log(prob.clamp_min(1e-10))).sum()); self.num_inst += prob.numel()

    def get(self):
        return self.name, (float("nan") if self.num_inst == 0 else math.exp(self.sum_metric / self.num_inst))


class PearsonCorrelation(EvalMetric):
    def __init__(self, name="pearsonr", **kw):
        super().__init__(name, **kw)

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).reshape(-1).double(), _t(p).reshape(-1).double()
            self.sum_metric += float(torch.corrcoef(torch.stack([l.to(p.device), p]))[0, 1]); self.num_inst += 1


class CustomMetric(EvalMetric):
    """Wraps ``feval(label_numpy, pred_numpy) -> float | (sum, count)`` (metric.py CustomMetric / np)."""

    def __init__(self, feval, name=None, allow_extra_outputs=False, **kw):
        self._feval = feval
        super().__init__(name or getattr(feval, "__name__", "custom"), **kw)

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            r = self._feval(_t(l).detach().cpu().numpy(), _t(p).detach().cpu().numpy())
            if isinstance(r, tuple):
                self.sum_metric += r[0]; self.num_inst += r[1]
            else:
                self.sum_metric += r; self.num_inst += 1


class NegativeLogLikelihood(EvalMetric):
    """Mean of ``-log p[label]`` (metric.py NegativeLogLikelihood :1050-1110)."""

    def __init__(self, eps=1e-12, name="nll-loss", **kw):
        super().__init__(name, **kw); self.eps = eps

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).reshape(-1).long(), _t(p)
            assert l.numel() == p.shape[0], "shape mismatch: %s vs. %s" % (tuple(l.shape), tuple(p.shape))
            prob = p[torch.arange(l.numel(), device=p.device), l.to(p.device)]
            self.sum_metric += float((-torch.log(prob + self.eps)).sum()); self.num_inst += l.numel()


class MCC(EvalMetric):
    """Matthews correlation coefficient of a binary classifier; ``average='macro'`` averages per-batch values, ``'micro'`` pools the
    confusion counts (metric.py MCC :650-750)."""

    def __init__(self, name="mcc", average="macro", **kw):
        self._average = average
        super().__init__(name, **kw)

    def reset(self):
        super().reset()
        self._c = [0.0, 0.0, 0.0, 0.0]                     # tp, fp, fn, tn

    @staticmethod
    def _mcc(tp, fp, fn, tn):
        den = math.sqrt(max((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn), 0.0))
        return (tp * tn - fp * fn) / den if den > 0 else 0.0

    def update(self, labels, preds):
        for l, p in zip(*_lists(labels, preds)):
            l, p = _t(l).reshape(-1).long(), _t(p)
            pred = p.argmax(dim=1) if p.dim() > 1 else (p > 0.5).long()
            l = l.to(pred.device)
            c = [float(((pred == 1) & (l == 1)).sum()), float(((pred == 1) & (l == 0)).sum()), float(((pred == 0) & (l == 1)).sum()), float(((pred == 0) & (l == 0)).sum())]
            if self._average == "macro":
                self.sum_metric += self._mcc(*c); self.num_inst += 1
            else:
                self._c = [a + b for a, b in zip(self._c, c)]
                self.sum_metric, self.num_inst = self._mcc(*self._c), 1


class Torch(Loss):
    """Dummy metric for outputs that already ARE a loss value (legacy torch criterions)."""

    def __init__(self, name="torch", **kw):
        super().__init__(name, **kw)


class Caffe(Torch):
    def __init__(self, name="caffe", **kw):
        super().__init__(name, **kw)


def check_label_shapes(labels, preds, wrap=False, shape=False):
    """Raise when the number (or, with ``shape=True``, the shapes) of labels and predictions differ; ``wrap`` puts single arrays in lists."""
    ls, ps = (labels.shape, preds.shape) if shape else (len(labels), len(preds))
    if ls != ps:
        raise ValueError("Shape of labels {} does not match shape of predictions {}".format(ls, ps))
    if wrap:
        labels = [labels] if hasattr(labels, "_t") else labels
        preds = [preds] if hasattr(preds, "_t") else preds
    return labels, preds


def np(numpy_feval, name=None, allow_extra_outputs=False):
    """Create a metric from a numpy function (``mx.metric.np``)."""
    return CustomMetric(numpy_feval, name, allow_extra_outputs)


_REG = {"mcc": MCC, "nll_loss": NegativeLogLikelihood, "nll-loss": NegativeLogLikelihood, "negativeloglikelihood": NegativeLogLikelihood, "torch": Torch, "caffe": Caffe,
        "compositeevalmetric": None, "topkaccuracy": TopKAccuracy, "crossentropy": CrossEntropy, "pearsoncorrelation": PearsonCorrelation, "f1": F1, "perplexity": Perplexity, "pearsonr": PearsonCorrelation, "acc": Accuracy, "accuracy": Accuracy, "top_k_accuracy": TopKAccuracy, "mae": MAE, "mse": MSE, "rmse": RMSE,
        "ce": CrossEntropy, "cross-entropy": CrossEntropy, "loss": Loss}


def create(metric, *args, **kwargs):
    if isinstance(metric, EvalMetric):
        return metric
    if callable(metric):
        return CustomMetric(metric, *args, **kwargs)
    if isinstance(metric, list):
        c = CompositeEvalMetric()
        for m in metric:
            c.add(create(m, *args, **kwargs))
        return c
    if metric.lower() == "compositeevalmetric":
        return CompositeEvalMetric(*args, **kwargs)
    return _REG[metric.lower()](*args, **kwargs)


def register(klass, name=None):
    """Register a user metric class under its lower-cased name (``mx.metric.register``)."""
    _REG[(name or klass.__name__).lower()] = klass
    return klass


def alias(*aliases):
    def deco(klass):
        for a in aliases:
            register(klass, a)
        return klass
    return deco
