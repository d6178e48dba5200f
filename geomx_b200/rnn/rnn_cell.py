"""``mx.rnn`` symbolic cells: build unrolled ``mx.sym`` graphs for ``mx.mod.Module`` / ``BucketingModule``.

Capability parity with ``python/mxnet/rnn/rnn_cell.py`` (RNNParams, BaseRNNCell.{begin_state, unroll, unpack_weights, pack_weights},
RNNCell / LSTMCell / GRUCell, FusedRNNCell + unfuse, SequentialRNNCell, BidirectionalCell, DropoutCell, ZoneoutCell, ResidualCell), with a
different construction: every gated cell is described by ONE table — gate names and a step function over the pre-activation slices — and a
single ``_GatedCell`` implements parameters, the step, and weight (un)packing for all of them; ``FusedRNNCell`` maps the same table onto the
flat parameter vector of the fused ``RNN`` operator (layout documented at ``mx.nd.RNN``).
"""
from __future__ import annotations

import numpy as np

from .. import ndarray as nd
from .. import symbol as sym
from ..base import MXNetError

__all__ = ["RNNParams", "BaseRNNCell", "RNNCell", "LSTMCell", "GRUCell", "FusedRNNCell", "SequentialRNNCell", "BidirectionalCell", "DropoutCell",
           "ModifierCell", "ZoneoutCell", "ResidualCell"]


class RNNParams:
    """Container of the (shared) parameter Variables of a cell: ``get('i2h_weight')`` -> ``Variable(prefix + 'i2h_weight')``, created once."""

    def __init__(self, prefix=""):
        self._prefix, self._params = prefix, {}

    def get(self, name, **kwargs):
        full = self._prefix + name
        if full not in self._params:
            self._params[full] = sym.Variable(full, **kwargs)
        return self._params[full]


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _split_inputs(length, inputs, layout, prefix):
    """``inputs``: one Symbol of layout ``layout`` ('NTC' / 'TNC') or a list of ``length`` step Symbols -> list of step Symbols [N, C]."""
    if isinstance(inputs, sym.Symbol):
        axis = layout.find("T")
        parts = sym.SliceChannel(inputs, num_outputs=length, axis=axis, squeeze_axis=1, name=prefix + "t_split")
        return [parts[i] for i in range(length)] if length > 1 else [parts[0] if hasattr(parts, "__getitem__") else parts]
    inputs = list(inputs)
    if len(inputs) != length:
        raise MXNetError("unroll: got %d step inputs for length %d" % (len(inputs), length))
    return inputs


def _merge_outputs(outputs, layout, merge):
    if not merge:
        return outputs
    axis = layout.find("T")
    return sym.Concat(*[sym.expand_dims(o, axis=axis) for o in outputs], dim=axis)


class BaseRNNCell:
    """One step of a recurrent network as a graph builder.  Sub-classes provide ``state_info`` and ``__call__(inputs, states)``."""

    def __init__(self, prefix="", params=None):
        self._own_params = params is None
        self._prefix = prefix
        self._params = RNNParams(prefix) if params is None else params
        self._modified = False
        self.reset()

    def reset(self):
        """Forget the step counter (call before building another graph with the same cell)."""
        self._init_counter = -1
        self._counter = -1
        for c in getattr(self, "_cells", []):
            c.reset()

    @property
    def params(self):
        self._own_params = False
        return self._params

    @property
    def state_info(self):
        raise NotImplementedError

    @property
    def state_shape(self):
        return [s["shape"] for s in self.state_info]

    @property
    def _gate_names(self):
        return ()

    def __call__(self, inputs, states):
        raise NotImplementedError

    def begin_state(self, func=None, **kwargs):
        """Initial states: ``func(name=..., **state_info, **kwargs)`` per state (default ``mx.sym.zeros``; pass ``mx.sym.Variable`` to feed them)."""
        if self._modified:
            raise MXNetError("After applying a modifier cell (e.g. DropoutCell) the base cell cannot be called directly; call the modifier cell instead.")
        func = func or sym.zeros
        batch_size = kwargs.pop("batch_size", 0)
        kwargs.pop("batch_ref", None)
        out = []
        for info in self.state_info:
            self._init_counter += 1
            kw = dict(kwargs)
            if func is sym.Variable:
                kw.pop("shape", None)
                out.append(sym.Variable("%sbegin_state_%d" % (self._prefix, self._init_counter), **kw))
            else:
                shape = tuple(batch_size if d == 0 else d for d in info["shape"])       # 0 = "the batch size": pass batch_size=N to fix it
                kw.setdefault("shape", shape)
                out.append(func(name="%sbegin_state_%d" % (self._prefix, self._init_counter), **kw))
        return out

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        """Apply the cell ``length`` times.  Returns ``(outputs, states)``: a list of step outputs (or one merged Symbol of ``layout`` when
        ``merge_outputs``) and the final states."""
        self.reset()
        steps = _split_inputs(length, inputs, layout, self._prefix)
        states = begin_state if begin_state is not None else self.begin_state(batch_ref=steps[0])
        outputs = []
        for x in steps:
            y, states = self(x, states)
            outputs.append(y)
        return _merge_outputs(outputs, layout, merge_outputs), states

    # ---- weights: fused per-layer matrices <-> one array per gate --------------------------------------------------------------------
    def unpack_weights(self, args):
        """``{prefix}i2h_weight`` [G*H, C] etc. -> one entry per gate (``{prefix}i2h{gate}_weight`` ...)."""
        args = dict(args)
        gates = self._gate_names
        if not gates:
            return args
        h = self._num_hidden
        for group in ("i2h", "h2h"):
            w = args.pop("%s%s_weight" % (self._prefix, group))
            b = args.pop("%s%s_bias" % (self._prefix, group))
            for j, gate in enumerate(gates):
                args["%s%s%s_weight" % (self._prefix, group, gate)] = w[j * h:(j + 1) * h].copy()
                args["%s%s%s_bias" % (self._prefix, group, gate)] = b[j * h:(j + 1) * h].copy()
        return args

    def pack_weights(self, args):
        args = dict(args)
        gates = self._gate_names
        if not gates:
            return args
        for group in ("i2h", "h2h"):
            ws = [args.pop("%s%s%s_weight" % (self._prefix, group, g)) for g in gates]
            bs = [args.pop("%s%s%s_bias" % (self._prefix, group, g)) for g in gates]
            args["%s%s_weight" % (self._prefix, group)] = nd.concat(*ws, dim=0)
            args["%s%s_bias" % (self._prefix, group)] = nd.concat(*bs, dim=0)
        return args


def _zeros_like_batch(x, hidden, name):
    """[N, hidden] zeros with the batch size of step input ``x`` (graphs stay shape-agnostic: no batch size is baked in)."""
    return sym.broadcast_mul(sym.zeros_like(sym.slice_axis(x, axis=1, begin=0, end=1)), sym.zeros(shape=(1, hidden)), name=name)


# ---- the gate tables: (gate names, number of states, step(pre-activation slices of i2h+h2h or of each, states) -> (output, new states)) -----
def _step_rnn(act):
    def step(i2h, h2h, states, name):
        out = sym.Activation(i2h[0] + h2h[0], act_type=act, name=name + "out")
        return out, [out]
    return step


def _step_lstm(i2h, h2h, states, name):
    g = [a + b for a, b in zip(i2h, h2h)]
    i = sym.Activation(g[0], act_type="sigmoid", name=name + "i")
    f = sym.Activation(g[1], act_type="sigmoid", name=name + "f")
    c_in = sym.Activation(g[2], act_type="tanh", name=name + "c")
    o = sym.Activation(g[3], act_type="sigmoid", name=name + "o")
    c = f * states[1] + i * c_in
    h = o * sym.Activation(c, act_type="tanh", name=name + "state")
    return h, [h, c]


def _step_gru(i2h, h2h, states, name):
    r = sym.Activation(i2h[0] + h2h[0], act_type="sigmoid", name=name + "r")
    z = sym.Activation(i2h[1] + h2h[1], act_type="sigmoid", name=name + "z")
    n = sym.Activation(i2h[2] + r * h2h[2], act_type="tanh", name=name + "h")
    h = (1.0 - z) * n + z * states[0]
    return h, [h]


class _GatedCell(BaseRNNCell):
    GATES, NSTATES = ("",), 1

    def __init__(self, num_hidden, prefix, params=None, i2h_bias_init=None):
        super().__init__(prefix=prefix, params=params)
        self._num_hidden = int(num_hidden)
        self._iW, self._iB = self.params.get("i2h_weight"), self.params.get("i2h_bias", **({"init": i2h_bias_init} if i2h_bias_init else {}))
        self._hW, self._hB = self.params.get("h2h_weight"), self.params.get("h2h_bias")

    @property
    def state_info(self):
        return [{"shape": (0, self._num_hidden), "__layout__": "NC"} for _ in range(self.NSTATES)]

    @property
    def _gate_names(self):
        return self.GATES

    def begin_state(self, func=None, batch_ref=None, **kwargs):
        if func is None and batch_ref is not None:           # zeros that follow the batch size of the data
            out = []
            for _ in self.state_info:
                self._init_counter += 1
                out.append(_zeros_like_batch(batch_ref, self._num_hidden, "%sbegin_state_%d" % (self._prefix, self._init_counter)))
            return out
        return super().begin_state(func=func, **kwargs)

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        G, H = len(self.GATES), self._num_hidden
        i2h = sym.FullyConnected(data=inputs, weight=self._iW, bias=self._iB, num_hidden=G * H, name=name + "i2h")
        h2h = sym.FullyConnected(data=states[0], weight=self._hW, bias=self._hB, num_hidden=G * H, name=name + "h2h")
        if G > 1:
            si = sym.SliceChannel(i2h, num_outputs=G, name=name + "i2h_slice")
            sh = sym.SliceChannel(h2h, num_outputs=G, name=name + "h2h_slice")
            i2h, h2h = [si[k] for k in range(G)], [sh[k] for k in range(G)]
        else:
            i2h, h2h = [i2h], [h2h]
        return self._step(i2h, h2h, states, name)


class RNNCell(_GatedCell):
    """Elman cell ``h' = act(W x + b + U h + c)``."""
    GATES, NSTATES = ("",), 1

    def __init__(self, num_hidden, activation="tanh", prefix="rnn_", params=None):
        super().__init__(num_hidden, prefix, params)
        self._step = _step_rnn(activation)


class LSTMCell(_GatedCell):
    """LSTM; gate order i, f, c, o; ``forget_bias`` is the initial value of the forget gate's i2h bias."""
    GATES, NSTATES = ("_i", "_f", "_c", "_o"), 2

    def __init__(self, num_hidden, prefix="lstm_", params=None, forget_bias=1.0):
        from .. import initializer
        super().__init__(num_hidden, prefix, params, i2h_bias_init=initializer.LSTMBias(forget_bias=forget_bias))
        self._step = _step_lstm


class GRUCell(_GatedCell):
    """GRU (cuDNN variant: the reset gate multiplies ``U_n h + c_n``); gate order r, z, o."""
    GATES, NSTATES = ("_r", "_z", "_o"), 1

    def __init__(self, num_hidden, prefix="gru_", params=None):
        super().__init__(num_hidden, prefix, params)
        self._step = _step_gru


_MODES = {"rnn_relu": (RNNCell, {"activation": "relu"}, 1), "rnn_tanh": (RNNCell, {"activation": "tanh"}, 1), "lstm": (LSTMCell, {}, 4), "gru": (GRUCell, {}, 3)}


class FusedRNNCell(BaseRNNCell):
    """A whole multi-layer (bidirectional) network as ONE ``RNN`` operator over a flat ``{prefix}parameters`` vector.  ``unroll`` only
    (no single-step ``__call__``); ``unfuse()`` gives the equivalent stack of step cells, ``unpack_weights`` / ``pack_weights`` convert
    between the flat vector and the stack's per-layer arrays."""

    def __init__(self, num_hidden, num_layers=1, mode="lstm", bidirectional=False, dropout=0.0, get_next_state=False, forget_bias=1.0, prefix=None, params=None):
        if mode not in _MODES:
            raise MXNetError("FusedRNNCell: unknown mode %r" % mode)
        super().__init__(prefix="%s_" % mode if prefix is None else prefix, params=params)
        self._num_hidden, self._num_layers, self._mode, self._bidirectional = int(num_hidden), int(num_layers), mode, bool(bidirectional)
        self._dropout, self._get_next_state, self._forget_bias = float(dropout), bool(get_next_state), forget_bias
        self._directions = ["l", "r"] if bidirectional else ["l"]
        from .. import initializer
        # the flat vector names its own initializer: blocks are filled by whatever initializer the caller passes to init_params, biases are
        # zero, LSTM forget-gate biases `forget_bias`
        self._parameter = self.params.get("parameters", init=initializer.FusedRNN(None, num_hidden, num_layers, mode, bidirectional, forget_bias))

    @property
    def state_info(self):
        b = self._num_layers * len(self._directions)
        n = 2 if self._mode == "lstm" else 1
        return [{"shape": (b, 0, self._num_hidden), "__layout__": "LNC"} for _ in range(n)]

    @property
    def _gate_names(self):
        return _MODES[self._mode][0].GATES

    def _layout(self, input_size):
        """[(name, shape)] in the order the fused vector stores them: all weights (layer by layer, direction by direction, i2h then h2h),
        then all biases in the same order."""
        G, H, D = _MODES[self._mode][2], self._num_hidden, len(self._directions)
        ws, bs = [], []
        for layer in range(self._num_layers):
            cin = input_size if layer == 0 else D * H
            for d in self._directions:
                p = "%s%s%d_" % (self._prefix, d, layer)
                ws += [(p + "i2h_weight", (G * H, cin)), (p + "h2h_weight", (G * H, H))]
                bs += [(p + "i2h_bias", (G * H,)), (p + "h2h_bias", (G * H,))]
        return ws + bs

    def _input_size(self, total):
        G, H, D, L = _MODES[self._mode][2], self._num_hidden, len(self._directions), self._num_layers
        rest = total - D * G * H * (2 + H) - (L - 1) * D * G * H * (D * H + H + 2)
        if rest <= 0 or rest % (D * G * H):
            raise MXNetError("FusedRNNCell: a parameter vector of %d elements does not fit this configuration" % total)
        return rest // (D * G * H)

    def unpack_weights(self, args):
        args = dict(args)
        flat = args.pop(self._prefix + "parameters")
        arr = flat.asnumpy().reshape(-1)
        pos = 0
        for name, shape in self._layout(self._input_size(arr.size)):
            n = int(np.prod(shape))
            args[name] = nd.array(arr[pos:pos + n].reshape(shape))
            pos += n
        # ... and further down to one array per gate, the format of the unfused stack
        for layer in range(self._num_layers):
            for d in self._directions:
                cell = _MODES[self._mode][0](self._num_hidden, prefix="%s%s%d_" % (self._prefix, d, layer), **_MODES[self._mode][1])
                args = cell.unpack_weights(args)
        return args

    def pack_weights(self, args):
        args = dict(args)
        for layer in range(self._num_layers):
            for d in self._directions:
                cell = _MODES[self._mode][0](self._num_hidden, prefix="%s%s%d_" % (self._prefix, d, layer), **_MODES[self._mode][1])
                args = cell.pack_weights(args)
        first = args["%sl0_i2h_weight" % self._prefix]
        chunks = [args.pop(name).asnumpy().reshape(-1) for name, _ in self._layout(first.shape[1])]
        args[self._prefix + "parameters"] = nd.array(np.concatenate(chunks))
        return args

    def __call__(self, inputs, states):
        raise MXNetError("FusedRNNCell cannot be stepped. Please use unroll")

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        if not isinstance(inputs, sym.Symbol):               # step list -> one TNC tensor
            inputs = sym.Concat(*[sym.expand_dims(x, axis=0) for x in inputs], dim=0)
            data = inputs
        else:
            data = sym.swapaxes(inputs, dim1=0, dim2=1) if layout == "NTC" else inputs
        if begin_state is None:
            ref = sym.slice_axis(data, axis=0, begin=0, end=1)                                  # [1, N, C]
            b = self._num_layers * len(self._directions)
            zero = sym.broadcast_mul(sym.zeros_like(sym.slice_axis(ref, axis=2, begin=0, end=1)), sym.zeros(shape=(b, 1, self._num_hidden)))
            begin_state = [zero for _ in self.state_info]
        states = list(begin_state)
        kw = {"state_cell": states[1]} if self._mode == "lstm" else {}
        rnn = sym.RNN(data, parameters=self._parameter, state=states[0], state_size=self._num_hidden, num_layers=self._num_layers,
                      bidirectional=self._bidirectional, p=self._dropout, state_outputs=self._get_next_state, mode=self._mode,
                      name=self._prefix + "rnn", **kw)
        if self._get_next_state:
            new_states = [rnn[1], rnn[2]] if self._mode == "lstm" else [rnn[1]]     # (indexing an output > 0 first marks the node multi-output)
            outputs = rnn[0]
        else:
            outputs, new_states = rnn, []
        if layout == "NTC":
            outputs = sym.swapaxes(outputs, dim1=0, dim2=1)
        if merge_outputs is False:
            axis = layout.find("T")
            parts = sym.SliceChannel(outputs, num_outputs=length, axis=axis, squeeze_axis=1)
            outputs = [parts[i] for i in range(length)]
        return outputs, new_states

    def unfuse(self):
        """The same network as a ``SequentialRNNCell`` of step cells (parameter names ``{prefix}{l|r}{layer}_...``)."""
        stack = SequentialRNNCell()
        cls, extra, _ = _MODES[self._mode]
        make = (lambda p: cls(self._num_hidden, prefix=p, forget_bias=self._forget_bias)) if self._mode == "lstm" else (lambda p: cls(self._num_hidden, prefix=p, **extra))
        for layer in range(self._num_layers):
            if self._bidirectional:
                stack.add(BidirectionalCell(make("%sl%d_" % (self._prefix, layer)), make("%sr%d_" % (self._prefix, layer)), output_prefix="%sbi_l%d_" % (self._prefix, layer)))
            else:
                stack.add(make("%sl%d_" % (self._prefix, layer)))
            if self._dropout > 0 and layer != self._num_layers - 1:
                stack.add(DropoutCell(self._dropout, prefix="%s_dropout%d_" % (self._prefix, layer)))
        return stack


class SequentialRNNCell(BaseRNNCell):
    """Cells stacked on top of each other: the output of one is the input of the next; states are concatenated in order."""

    def __init__(self, params=None):
        self._cells = []
        super().__init__(prefix="", params=params)
        self._override_cell_params = params is not None

    def add(self, cell):
        self._cells.append(cell)
        if self._override_cell_params:
            if not cell._own_params:
                raise MXNetError("Either specify params for SequentialRNNCell or child cells, not both.")
            cell.params._params.update(self.params._params)
        self.params._params.update(cell.params._params)

    @property
    def state_info(self):
        return [s for c in self._cells for s in c.state_info]

    def begin_state(self, **kwargs):
        return [s for c in self._cells for s in c.begin_state(**kwargs)]

    def unpack_weights(self, args):
        for c in self._cells:
            args = c.unpack_weights(args)
        return args

    def pack_weights(self, args):
        for c in self._cells:
            args = c.pack_weights(args)
        return args

    def __call__(self, inputs, states):
        self._counter += 1
        nxt, pos = [], 0
        for c in self._cells:
            if isinstance(c, BidirectionalCell):
                raise MXNetError("BidirectionalCell cannot be stepped: use unroll")
            n = len(c.state_info)
            inputs, st = c(inputs, states[pos:pos + n])
            pos += n
            nxt += st
        return inputs, nxt

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        steps = _split_inputs(length, inputs, layout, "seq_")
        if begin_state is None:
            begin_state = [s for c in self._cells for s in (c.begin_state(batch_ref=steps[0]) if isinstance(c, (_GatedCell, ModifierCell, BidirectionalCell)) else c.begin_state())]
        pos, nxt, cur = 0, [], steps
        for i, c in enumerate(self._cells):
            n = len(c.state_info)
            cur, st = c.unroll(length, cur, begin_state=begin_state[pos:pos + n], layout=layout, merge_outputs=None if i < len(self._cells) - 1 else merge_outputs)
            if isinstance(cur, sym.Symbol) and i < len(self._cells) - 1:
                cur = _split_inputs(length, cur, layout, "seq%d_" % i)
            pos += n
            nxt += st
        return cur, nxt


class DropoutCell(BaseRNNCell):
    """Dropout on the data path between stacked cells (no state)."""

    def __init__(self, dropout, prefix="dropout_", params=None):
        super().__init__(prefix, params)
        self.dropout = float(dropout)

    @property
    def state_info(self):
        return []

    def begin_state(self, **kwargs):
        return []

    def __call__(self, inputs, states):
        return (sym.Dropout(inputs, p=self.dropout) if self.dropout > 0 else inputs), states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        if isinstance(inputs, sym.Symbol):
            return self(inputs, [])
        outs = [self(x, [])[0] for x in inputs]
        return _merge_outputs(outs, layout, merge_outputs), []


class ModifierCell(BaseRNNCell):
    """A cell that wraps another one and changes what goes in or comes out; the wrapped cell must then be used through the wrapper only."""

    def __init__(self, base_cell):
        base_cell._modified = True
        super().__init__()
        self.base_cell = base_cell

    @property
    def params(self):
        self._own_params = False
        return self.base_cell.params

    @property
    def state_info(self):
        return self.base_cell.state_info

    def begin_state(self, func=None, **kwargs):
        self.base_cell._modified = False
        try:
            return self.base_cell.begin_state(func=func, **kwargs)
        finally:
            self.base_cell._modified = True

    def unpack_weights(self, args):
        return self.base_cell.unpack_weights(args)

    def pack_weights(self, args):
        return self.base_cell.pack_weights(args)

    def reset(self):
        super().reset()
        if hasattr(self, "base_cell"):
            self.base_cell.reset()


class ZoneoutCell(ModifierCell):
    """Zoneout (Krueger et al.): with probability p an output / state keeps its previous value instead of the new one (training only)."""

    def __init__(self, base_cell, zoneout_outputs=0.0, zoneout_states=0.0):
        if isinstance(base_cell, (FusedRNNCell, BidirectionalCell)):
            raise MXNetError("ZoneoutCell needs a cell that can be stepped")
        super().__init__(base_cell)
        self.zoneout_outputs, self.zoneout_states = float(zoneout_outputs), float(zoneout_states)
        self.prev_output = None

    def reset(self):
        super().reset()
        self.prev_output = None

    def __call__(self, inputs, states):
        out, nxt = self.base_cell(inputs, states)

        def mix(p, new, old):
            if p <= 0:
                return new
            mask = sym.Dropout(sym.ones_like(new), p=p)             # 0 with probability p (scaled ones elsewhere): where() only tests != 0
            return sym.where(mask, new, old)
        prev = self.prev_output if self.prev_output is not None else sym.zeros_like(out)
        output = mix(self.zoneout_outputs, out, prev)
        nxt = [mix(self.zoneout_states, n, o) for n, o in zip(nxt, states)]
        self.prev_output = output
        return output, nxt


class ResidualCell(ModifierCell):
    """``output = cell(input) + input`` (He et al.); input and hidden sizes must agree."""

    def __call__(self, inputs, states):
        out, states = self.base_cell(inputs, states)
        return out + inputs, states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        steps = _split_inputs(length, inputs, layout, "res_")
        self.base_cell._modified = False
        try:
            outs, states = self.base_cell.unroll(length, steps, begin_state=begin_state, layout=layout, merge_outputs=False)
        finally:
            self.base_cell._modified = True
        outs = [o + x for o, x in zip(outs, steps)]
        return _merge_outputs(outs, layout, merge_outputs), states


class BidirectionalCell(BaseRNNCell):
    """One cell over the sequence, another over the reversed sequence, outputs concatenated step by step.  ``unroll`` only."""

    def __init__(self, l_cell, r_cell, params=None, output_prefix="bi_"):
        self._cells = [l_cell, r_cell]
        super().__init__("", params=params)
        self._output_prefix = output_prefix
        self._override_cell_params = params is not None
        if self._override_cell_params:
            if not (l_cell._own_params and r_cell._own_params):
                raise MXNetError("Either specify params for BidirectionalCell or child cells, not both.")
            l_cell.params._params.update(self.params._params); r_cell.params._params.update(self.params._params)
        self.params._params.update(l_cell.params._params); self.params._params.update(r_cell.params._params)

    @property
    def state_info(self):
        return [s for c in self._cells for s in c.state_info]

    def begin_state(self, **kwargs):
        return [s for c in self._cells for s in c.begin_state(**kwargs)]

    def unpack_weights(self, args):
        for c in self._cells:
            args = c.unpack_weights(args)
        return args

    def pack_weights(self, args):
        for c in self._cells:
            args = c.pack_weights(args)
        return args

    def __call__(self, inputs, states):
        raise MXNetError("Bidirectional cannot be stepped. Please use unroll")

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        steps = _split_inputs(length, inputs, layout, self._output_prefix)
        l, r = self._cells
        if begin_state is None:
            begin_state = l.begin_state(batch_ref=steps[0]) + r.begin_state(batch_ref=steps[0])
        nl = len(l.state_info)
        lo, ls = l.unroll(length, steps, begin_state=begin_state[:nl], layout=layout, merge_outputs=False)
        ro, rs = r.unroll(length, list(reversed(steps)), begin_state=begin_state[nl:], layout=layout, merge_outputs=False)
        outs = [sym.Concat(a, b, dim=1, name="%st%d" % (self._output_prefix, i)) for i, (a, b) in enumerate(zip(lo, reversed(ro)))]
        return _merge_outputs(outs, layout, merge_outputs), ls + rs
