"""``mx.visualization.print_summary`` (parity: python/mxnet/visualization.py:36-200): layer table of a Symbol with output shapes and
parameter counts.  (``plot_network`` needs graphviz, which is not part of this environment.)"""
from __future__ import annotations

import numpy as np

__all__ = ["print_summary"]


def print_summary(symbol, shape=None, line_length=100, positions=(.44, .64, .74, 1.)):
    shapes = {}
    if shape:
        arg_shapes, _, aux_shapes = symbol.infer_shape(**shape)
        shapes.update(dict(zip(symbol.list_arguments(), arg_shapes)))
        shapes.update(dict(zip(symbol.list_auxiliary_states(), aux_shapes)))
    cols = [int(line_length * p) for p in positions]
    heads = ["Layer (type)", "Output Shape", "Param #", "Previous Layer"]

    def row(fields):
        line = ""
        for f, c in zip(fields, cols):
            line = (line + str(f))[:c].ljust(c)
        return line
    lines = ["_" * line_length, row(heads), "=" * line_length]
    total = 0
    for node in symbol._topo():
        if node.op in ("null", "_group"):
            continue
        out_shape = ""
        if shape:
            try:
                _, o, _ = node.infer_shape(**{k: v for k, v in shape.items()})
                out_shape = "x".join(str(d) for d in o[0][1:])
            except Exception:
                out_shape = "?"
        params = 0
        for inp in node.inputs[1:] + node.aux:
            if inp.op == "null" and inp.name in shapes and shapes[inp.name] is not None and not inp.name.endswith("label"):
                params += int(np.prod(shapes[inp.name]))
        total += params
        prev = ", ".join(i.name for i in node.inputs[:1] if i.op != "null" or i.name in (shape or {}))
        lines += [row(["%s(%s)" % (node.name, node.op), out_shape, params, prev]), "_" * line_length]
    lines += ["Total params: %d" % total, "_" * line_length]
    text = "\n".join(lines)
    print(text)
    return total
