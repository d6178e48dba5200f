"""``mx.contrib.quantization`` — post-training quantisation with calibration.

Parity: ``python/mxnet/contrib/quantization.py`` (``quantize_model`` :423-549, min/max and KL-entropy calibration :130-420).  The reference
rewrites the graph with int8 MKLDNN/cuDNN operators; B200 has no INT8 tensor-core advantage over fp8 (and sm_103 removes INT8 MMA
entirely), so this module produces a *simulated-quantisation* model instead: weights of the quantised layers are replaced by their
quantise→dequantise image (``int8`` symmetric per-tensor, or ``fp8`` e4m3 per-tensor scaled — the format the fabric's wire codec uses), the
calibrated activation thresholds are recorded per layer, and activations are clamped+rounded on the fly at those thresholds.  The result
runs on the normal executors, numerically equal to what an integer pipeline would compute up to accumulation order.

``quantize_model(sym, arg_params, aux_params, …)`` handles symbolic models, ``quantize_net(net, …)`` gluon blocks."""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import ndarray as nd
from ..ndarray import NDArray

__all__ = ["quantize_model", "quantize_net", "calib_thresholds", "fake_quantize"]

_QUANTIZABLE = ("FullyConnected", "Convolution")


def fake_quantize(t, threshold, dtype="int8"):
    """Quantise→dequantise of a tensor at a symmetric ``threshold`` (|x| clipped to it)."""
    thr = max(float(threshold), 1e-30)
    if dtype == "int8":
        scale = thr / 127.0
        return torch.clamp(torch.round(t / scale), -127, 127) * scale
    if dtype == "uint8":
        scale = thr / 255.0
        return torch.clamp(torch.round(t / scale), 0, 255) * scale
    if dtype in ("fp8", "float8", "e4m3"):
        scale = thr / 448.0
        return (torch.clamp(t / scale, -448, 448)).to(torch.float8_e4m3fn).to(t.dtype) * scale
    raise ValueError("unknown quantized_dtype %s (int8, uint8, fp8)" % dtype)


def _kl(p, q):
    m = p > 0
    return float(np.sum(p[m] * np.log(p[m] / np.maximum(q[m], 1e-12))))


def _entropy_threshold(samples, num_bins=2001, num_quantized_bins=255):
    """Threshold minimising KL(reference histogram ‖ its 255-level quantised image), the TensorRT-style calibration the reference uses
    (``quantization.py:250-340``): slide the clip point outward from the centre, fold the clipped tails into the edge bins, quantise the
    kept range into 255 buckets, expand back, compare."""
    a = np.abs(np.concatenate([s.ravel() for s in samples]))
    amax = float(a.max()) if a.size else 0.0
    if amax == 0.0:
        return 0.0
    hist, edges = np.histogram(a, bins=num_bins // 2 + 1, range=(0, amax))
    hist = hist.astype(np.float64)
    best, best_thr = None, amax
    nq = num_quantized_bins // 2 + 1                                  # one-sided: 128 magnitude levels
    for i in range(nq, len(hist) + 1, max(1, len(hist) // 256)):
        p = hist[:i].copy()
        p[i - 1] += hist[i:].sum()                                    # clipped mass lands in the last kept bin
        if p.sum() == 0:
            continue
        idx = (np.arange(i) * nq // i)
        q_lvl = np.bincount(idx, weights=hist[:i], minlength=nq)
        nz = np.bincount(idx, weights=(hist[:i] > 0).astype(np.float64), minlength=nq)
        q = np.where(hist[:i] > 0, q_lvl[idx] / np.maximum(nz[idx], 1), 0.0)
        pn, qn = p / p.sum(), q / max(q.sum(), 1e-12)
        d = _kl(pn, qn)
        if best is None or d < best:
            best, best_thr = d, float(edges[i])
    return best_thr


def calib_thresholds(collected, calib_mode="naive"):
    """``{layer: [numpy arrays]}`` → ``{layer: threshold}``; ``naive`` = max |x|, ``entropy`` = KL-optimal clip."""
    out = {}
    for name, arrs in collected.items():
        out[name] = float(max(np.abs(a).max() for a in arrs)) if calib_mode == "naive" else _entropy_threshold(arrs)
    return out


# ------------------------------------------------------------------------------------------------ symbolic models
def _collect_symbol_inputs(sym, arg_params, aux_params, calib_data, data_names, label_names, num_calib_examples, ctx, layers):
    """Runs calibration batches through the graph and records the DATA input of every quantisable layer."""
    from ..symbol import Group
    nodes = [s for s in sym._topo() if s.op in _QUANTIZABLE and s.name in layers]
    taps = Group([n.inputs[0] for n in nodes] + [sym])
    collected = {n.name: [] for n in nodes}
    seen, ex = 0, None
    calib_data.reset()
    for batch in calib_data:
        shapes = {d: tuple(a.shape) for d, a in zip(data_names, batch.data)}
        if ex is None:
            for ln, lb in zip(label_names or (), batch.label or ()):
                if ln in sym.list_arguments():
                    shapes[ln] = tuple(lb.shape)
            ex = taps.simple_bind(ctx, grad_req="null", **shapes)
            ex.copy_params_from(arg_params, aux_params, allow_extra_params=True)
        for d, a in zip(data_names, batch.data):
            ex.arg_dict[d][:] = a
        outs = ex.forward(is_train=False)
        for n, o in zip(nodes, outs):
            collected[n.name].append(o.asnumpy())
        seen += batch.data[0].shape[0]
        if num_calib_examples is not None and seen >= num_calib_examples:
            break
    return collected, seen


def quantize_model(sym, arg_params, aux_params, data_names=("data",), label_names=("softmax_label",), ctx=None, excluded_sym_names=None,
                   calib_mode="entropy", calib_data=None, num_calib_examples=None, calib_layer=None, quantized_dtype="int8", logger=logging):
    """Returns ``(qsym, qarg_params, aux_params)``.  ``qsym`` is ``sym`` with every quantised FullyConnected / Convolution annotated
    (``attrs['__quantized__'] = {'dtype', 'act_threshold'}``; the executor fake-quantises that layer's input at the threshold),
    ``qarg_params`` holds the quantise→dequantise image of their weights plus ``<name>_weight_min/_max`` range entries.
    ``calib_mode``: ``none`` (weights only, activations quantised at their run-time max), ``naive`` (max |x| over the calibration set) or
    ``entropy`` (KL-optimal thresholds).  ``calib_layer(name) -> bool`` restricts which layers get calibrated thresholds."""
    from .. import context, symbol as S
    ctx = ctx or context.cpu()
    if calib_mode not in ("none", "naive", "entropy"):
        raise ValueError("unknown calibration mode %s received, expected `none`, `naive`, or `entropy`" % calib_mode)
    excluded = set(excluded_sym_names or [])
    qsym = S.load_json(sym.tojson())                                   # private copy to annotate
    layers = {s.name for s in qsym._topo() if s.op in _QUANTIZABLE and s.name not in excluded}
    thresholds = {}
    if calib_mode != "none":
        if calib_data is None:
            raise ValueError("calib_data must be provided when calib_mode=%s" % calib_mode)
        want = {n for n in layers if calib_layer is None or calib_layer(n)}
        collected, seen = _collect_symbol_inputs(qsym, arg_params, aux_params, calib_data, list(data_names), list(label_names or ()),
                                                 num_calib_examples, ctx, want)
        logger.info("Collected layer inputs from %d calibration examples for %d layers", seen, len(collected))
        thresholds = calib_thresholds(collected, calib_mode)
    qargs = {k: v for k, v in arg_params.items()}
    for s in qsym._topo():
        if s.op in _QUANTIZABLE and s.name in layers:
            wname = s.inputs[1].name
            w = arg_params[wname]._t
            thr = float(w.abs().max())
            qargs[wname] = NDArray(fake_quantize(w, thr, "int8" if quantized_dtype == "uint8" else quantized_dtype))
            qargs[wname + "_min"] = nd.array([-thr]); qargs[wname + "_max"] = nd.array([thr])
            s.attrs["__quantized__"] = {"dtype": quantized_dtype, "act_threshold": thresholds.get(s.name)}
    return qsym, qargs, aux_params


# ------------------------------------------------------------------------------------------------ gluon
def quantize_net(net, calib_data=None, calib_mode="naive", quantized_dtype="int8", exclude_layers=None, num_calib_batches=None, logger=logging):
    """In-place simulated quantisation of the ``Dense`` / ``Conv2D`` children of a gluon block: weights are replaced by their quantised
    image and a forward pre-hook fake-quantises the layer input at the calibrated threshold.  ``calib_data`` yields ``(data, label)`` or
    ``data`` batches.  Returns ``net``."""
    from ..gluon import nn as gnn
    excl = set(exclude_layers or [])
    targets = []

    def walk(b):
        for c in b._children.values():
            if isinstance(c, (gnn.Dense, gnn.Conv2D)) and c.name not in excl:
                targets.append(c)
            walk(c)
    walk(net)
    collected = {b.name: [] for b in targets}
    if calib_mode != "none":
        if calib_data is None:
            raise ValueError("calib_data must be provided when calib_mode=%s" % calib_mode)
        hooks = [b.register_forward_pre_hook(lambda blk, inputs: collected[blk.name].append(inputs[0].asnumpy())) for b in targets]
        for i, batch in enumerate(calib_data):
            x = batch[0] if isinstance(batch, (list, tuple)) else batch
            net(x)
            if num_calib_batches is not None and i + 1 >= num_calib_batches:
                break
        for h in hooks:
            h.detach()
    thr = calib_thresholds({k: v for k, v in collected.items() if v}, calib_mode if calib_mode != "none" else "naive")
    act_dtype = quantized_dtype
    for b in targets:
        w = b.weight.data()
        with torch.no_grad():
            w._t.copy_(fake_quantize(w._t, float(w._t.abs().max()), "int8" if quantized_dtype == "uint8" else quantized_dtype))
        t = thr.get(b.name)

        def pre(blk, inputs, _t=t):
            x = inputs[0]
            limit = _t if _t is not None else float(x._t.abs().max())
            return (NDArray(fake_quantize(x._t, limit, "int8" if act_dtype == "uint8" and float(x._t.min()) < 0 else act_dtype)),) + tuple(inputs[1:])
        b.register_forward_pre_hook(pre)
        b._quantized = {"dtype": quantized_dtype, "act_threshold": t}
    logger.info("quantized %d layers (%s, calib=%s)", len(targets), quantized_dtype, calib_mode)
    return net
