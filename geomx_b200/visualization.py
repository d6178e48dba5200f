"""``mx.visualization.print_summary`` (parity: python/mxnet/visualization.py:36-200): layer table of a Symbol with output shapes and
parameter counts.  (``plot_network`` needs graphviz, which is not part of this environment.)"""
from __future__ import annotations

import numpy as np

__all__ = ["print_summary", "plot_network"]


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


class _Dot:
    """Minimal stand-in for ``graphviz.Digraph`` when the package is absent: collects DOT text (``.source``, ``.save(path)``)."""

    def __init__(self, name):
        self.name, self._lines = name, []

    def node(self, name, label, **attrs):
        a = ", ".join('%s="%s"' % (k, v) for k, v in dict(label=label, **attrs).items())
        self._lines.append('  "%s" [%s];' % (name, a))

    def edge(self, tail, head, label=""):
        self._lines.append('  "%s" -> "%s" [label="%s"];' % (tail, head, label))

    @property
    def source(self):
        return "digraph %s {\n%s\n}\n" % (self.name, "\n".join(self._lines))

    def save(self, filename):
        with open(filename, "w") as f:
            f.write(self.source)
        return filename


def plot_network(symbol, title="plot", save_format="pdf", shape=None, dtype=None, node_attrs=None, hide_weights=True):
    """Graph of a Symbol as a ``graphviz.Digraph`` (or, without the graphviz package, an object with the same ``node`` / ``edge`` /
    ``source`` surface holding DOT text).  Edges are labelled with tensor shapes when ``shape`` is given (visualization.py:210-400)."""
    try:
        from graphviz import Digraph
        dot = Digraph(name=title, format=save_format)
    except ImportError:
        dot = _Dot(title)
    shapes = {}
    if shape is not None:
        internals = symbol.get_internals()
        _, outs, _ = internals.infer_shape(**shape)
        shapes = dict(zip(internals.list_outputs(), outs))
    palette = {"FullyConnected": "#fb8072", "Convolution": "#fb8072", "Activation": "#ffffb3", "Pooling": "#80b1d3", "BatchNorm": "#bebada",
               "Flatten": "#fdb462", "SoftmaxOutput": "#b3de69"}
    hidden = set()
    for s in symbol._topo():
        if s.op == "_group":
            continue
        if s.op == "null":
            if hide_weights and any(s.name.endswith(x) for x in ("_weight", "_bias", "_gamma", "_beta", "_moving_mean", "_moving_var")):
                hidden.add(id(s)); continue
            dot.node(s.name, s.name, shape="oval", style="filled", fillcolor="#8dd3c7")
        else:
            a = s.attrs
            label = s.op
            if s.op == "Convolution":
                label = "Convolution\n%s/%s, %d" % ("x".join(map(str, a["kernel"])), "x".join(map(str, a["stride"])), a["num_filter"])
            elif s.op == "FullyConnected":
                label = "FullyConnected\n%d" % a["num_hidden"]
            elif s.op == "Activation":
                label = "Activation\n%s" % a["act_type"]
            elif s.op == "Pooling":
                label = "Pooling\n%s, %s" % (a["pool_type"], "x".join(map(str, a["kernel"])))
            dot.node(s.name, label, shape="box", style="filled", fillcolor=palette.get(s.op, "#fccde5"), **(node_attrs or {}))
            for i in s.inputs:
                if id(i) in hidden:
                    continue
                key = i.name if i.op == "null" else i.name + "_output"
                dot.edge(i.name, s.name, "x".join(map(str, shapes[key][1:])) if key in shapes else "")
    return dot
