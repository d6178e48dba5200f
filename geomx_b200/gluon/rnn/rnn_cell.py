"""Recurrent cells.  Gate order follows MXNet (LSTM: i, f, g(c~), o; GRU: r, z, n) so that ``.params`` files interchange."""
from __future__ import annotations

import torch

from ... import ndarray as nd
from ...ndarray import NDArray
from ...ops import functional as OF
from ..block import HybridBlock

__all__ = ["RecurrentCell", "RNNCell", "LSTMCell", "GRUCell", "SequentialRNNCell", "ModifierCell", "DropoutCell", "ZoneoutCell", "ResidualCell",
           "BidirectionalCell", "HybridRecurrentCell", "HybridSequentialRNNCell"]


class RecurrentCell(HybridBlock):
    _gates = 1

    def __init__(self, hidden_size, input_size=0, i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer="zeros",
                 h2h_bias_initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        self._hidden_size, self._input_size = hidden_size, input_size
        G = self._gates * hidden_size
        with self.name_scope():
            self.i2h_weight = self.params.get("i2h_weight", shape=(G, input_size), init=i2h_weight_initializer, allow_deferred_init=True)
            self.h2h_weight = self.params.get("h2h_weight", shape=(G, hidden_size), init=h2h_weight_initializer, allow_deferred_init=True)
            from ..nn.basic_layers import _bias_init
            self.i2h_bias = self.params.get("i2h_bias", shape=(G,), init=_bias_init(i2h_bias_initializer), allow_deferred_init=True)
            self.h2h_bias = self.params.get("h2h_bias", shape=(G,), init=_bias_init(h2h_bias_initializer), allow_deferred_init=True)

    def _infer(self, x, *a):
        self.i2h_weight.shape = (self._gates * self._hidden_size, x.shape[-1])

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size, self._hidden_size), "__layout__": "NC"}]

    def reset(self):
        """Forget the per-sequence bookkeeping (step counters used to name states) of this cell and its children; call between sequences
        when cells are stepped by hand (rnn_cell.py RecurrentCell.reset)."""
        self._init_counter = -1
        self._counter = -1
        for c in self._children.values():
            if hasattr(c, "reset"):
                c.reset()

    def begin_state(self, batch_size=0, func=None, ctx=None, **kwargs):
        func = func or nd.zeros
        return [func(info["shape"], ctx=ctx) for info in self.state_info(batch_size)]

    def __call__(self, inputs, states):
        return super().__call__(inputs, states)

    def forward(self, inputs, states):
        self._deferred_infer(inputs)
        p = [self.i2h_weight.data(inputs.context), self.h2h_weight.data(inputs.context), self.i2h_bias.data(inputs.context), self.h2h_bias.data(inputs.context)]
        return self._step(inputs, states, *p)

    def _deferred_infer(self, x):
        if self.i2h_weight.shape[1] == 0 or self.i2h_weight._data is None:
            self._infer(x)
            for prm in (self.i2h_weight, self.h2h_weight, self.i2h_bias, self.h2h_bias):
                prm._finish_deferred_init() if hasattr(prm, "_finish_deferred_init") else None

    def _gates_pre(self, x, h, wi, wh, bi, bh):
        return OF.dense(x._t, wi._t, bi._t, None, False) + OF.dense(h._t, wh._t, bh._t, None, False)

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        axis = layout.find("T")
        seq = [NDArray(t) for t in inputs._t.unbind(axis)] if isinstance(inputs, NDArray) else list(inputs)
        batch = seq[0].shape[0]
        states = begin_state or self.begin_state(batch, ctx=seq[0].context)
        outs = []
        for t in range(length):
            o, states = self(seq[t], states)
            outs.append(o)
        if merge_outputs or merge_outputs is None and isinstance(inputs, NDArray):
            return NDArray(torch.stack([o._t for o in outs], dim=axis)), states
        return outs, states


class RNNCell(RecurrentCell):
    def __init__(self, hidden_size, activation="tanh", **kwargs):
        super().__init__(hidden_size, **kwargs)
        self._activation = activation

    def _step(self, x, states, wi, wh, bi, bh):
        h = NDArray(OF.activation(self._gates_pre(x, states[0], wi, wh, bi, bh), self._activation))
        return h, [h]


class LSTMCell(RecurrentCell):
    _gates = 4

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size, self._hidden_size), "__layout__": "NC"}] * 2

    def _step(self, x, states, wi, wh, bi, bh):
        i, f, g, o = self._gates_pre(x, states[0], wi, wh, bi, bh).chunk(4, dim=-1)
        c = torch.sigmoid(f) * states[1]._t + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        return NDArray(h), [NDArray(h), NDArray(c)]


class GRUCell(RecurrentCell):
    _gates = 3

    def _step(self, x, states, wi, wh, bi, bh):
        gi = OF.dense(x._t, wi._t, bi._t, None, False).chunk(3, dim=-1)
        gh = OF.dense(states[0]._t, wh._t, bh._t, None, False).chunk(3, dim=-1)
        r = torch.sigmoid(gi[0] + gh[0]); z = torch.sigmoid(gi[1] + gh[1])
        n = torch.tanh(gi[2] + r * gh[2])
        h = (1 - z) * n + z * states[0]._t
        return NDArray(h), [NDArray(h)]


class SequentialRNNCell(RecurrentCell):
    def __init__(self, **kwargs):
        HybridBlock.__init__(self, **kwargs)
        self._cells = []

    def add(self, cell):
        self._cells.append(cell)
        self.register_child(cell)

    def state_info(self, batch_size=0):
        return [s for c in self._cells for s in c.state_info(batch_size)]

    def begin_state(self, batch_size=0, **kw):
        return [s for c in self._cells for s in c.begin_state(batch_size, **kw)]

    def forward(self, inputs, states):
        nxt, pos = [], 0
        for c in self._cells:
            n = len(c.state_info())
            inputs, st = c(inputs, states[pos:pos + n])
            nxt += st; pos += n
        return inputs, nxt


class DropoutCell(RecurrentCell):
    """Dropout on the cell input; stateless (``gluon/rnn/rnn_cell.py:700-750``)."""

    def __init__(self, rate, axes=(), **kwargs):
        HybridBlock.__init__(self, **kwargs); self._rate = rate

    def state_info(self, batch_size=0):
        return []

    def forward(self, inputs, states):
        from ... import autograd
        if self._rate > 0 and autograd.is_training():
            inputs = NDArray(torch.nn.functional.dropout(inputs._t, self._rate, True))
        return inputs, states


class ModifierCell(RecurrentCell):
    """Base of cells that wrap another cell and share its parameters / state layout (``rnn_cell.py:760-800``)."""

    def __init__(self, base_cell, **kwargs):
        HybridBlock.__init__(self, **kwargs)
        self.base_cell = base_cell
        self.register_child(base_cell)

    def state_info(self, batch_size=0):
        return self.base_cell.state_info(batch_size)

    def begin_state(self, batch_size=0, **kw):
        return self.base_cell.begin_state(batch_size, **kw)


class ZoneoutCell(ModifierCell):
    """Zoneout: randomly keep the previous output / states (``rnn_cell.py:803-860``)."""

    def __init__(self, base_cell, zoneout_outputs=0.0, zoneout_states=0.0, **kwargs):
        super().__init__(base_cell, **kwargs)
        self._zo, self._zs, self._prev = zoneout_outputs, zoneout_states, None

    def reset(self):
        self._prev = None

    def forward(self, inputs, states):
        from ... import autograd
        out, nxt = self.base_cell(inputs, states)
        if not autograd.is_training():
            self._prev = out
            return out, nxt
        def mix(p, new, old):
            if p <= 0 or old is None:
                return new
            keep = (torch.rand_like(new._t) < p).to(new._t.dtype)
            return NDArray(keep * old._t + (1 - keep) * new._t)
        prev = self._prev if self._prev is not None else NDArray(torch.zeros_like(out._t))
        out = mix(self._zo, out, prev)
        nxt = [mix(self._zs, n, o) for n, o in zip(nxt, states)]
        self._prev = out
        return out, nxt


class ResidualCell(ModifierCell):
    """``output = base(input) + input`` (``rnn_cell.py:863-900``)."""

    def forward(self, inputs, states):
        out, nxt = self.base_cell(inputs, states)
        return NDArray(out._t + inputs._t), nxt


class BidirectionalCell(RecurrentCell):
    """Runs one cell forward and one backward over a sequence and concatenates their outputs; only ``unroll`` is defined
    (``rnn_cell.py:903-1000``)."""

    def __init__(self, l_cell, r_cell, output_prefix="bi_", **kwargs):
        HybridBlock.__init__(self, **kwargs)
        self._cells = [l_cell, r_cell]
        self.register_child(l_cell); self.register_child(r_cell)

    def state_info(self, batch_size=0):
        return [s for c in self._cells for s in c.state_info(batch_size)]

    def begin_state(self, batch_size=0, **kw):
        return [s for c in self._cells for s in c.begin_state(batch_size, **kw)]

    def forward(self, inputs, states):
        raise NotImplementedError("BidirectionalCell cannot be stepped; use unroll")

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        axis = layout.find("T")
        seq = [NDArray(t) for t in inputs._t.unbind(axis)] if isinstance(inputs, NDArray) else list(inputs)
        batch = seq[0].shape[0]
        states = begin_state or self.begin_state(batch, ctx=seq[0].context)
        nl = len(self._cells[0].state_info())
        lo, ls = self._cells[0].unroll(length, seq, states[:nl], layout, merge_outputs=False)
        ro, rs = self._cells[1].unroll(length, seq[::-1], states[nl:], layout, merge_outputs=False)
        outs = [NDArray(torch.cat([a._t, b._t], dim=-1)) for a, b in zip(lo, ro[::-1])]
        if merge_outputs or merge_outputs is None and isinstance(inputs, NDArray):
            return NDArray(torch.stack([o._t for o in outs], dim=axis)), ls + rs
        return outs, ls + rs


# every cell here is a HybridBlock already; the reference keeps separate names for the hybridizable bases (rnn_cell.py:300-330, :690)
HybridRecurrentCell = RecurrentCell
HybridSequentialRNNCell = SequentialRNNCell
