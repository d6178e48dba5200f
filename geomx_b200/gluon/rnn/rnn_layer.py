"""Multi-layer (bi)directional recurrent layers composed of the cells (the reference dispatches to cuDNN's fused RNN; here the time loop is
explicit and every step's GEMMs go through ``ops.functional.dense``)."""
from __future__ import annotations

import torch

from ... import ndarray as nd
from ...ndarray import NDArray
from ..block import HybridBlock
from .rnn_cell import GRUCell, LSTMCell, RNNCell

__all__ = ["RNN", "LSTM", "GRU"]


class _RNNLayer(HybridBlock):
    _cell = None

    def __init__(self, hidden_size, num_layers=1, layout="TNC", dropout=0.0, bidirectional=False, input_size=0, **kwargs):
        cell_kw = {k: kwargs.pop(k) for k in list(kwargs) if k in ("activation",)}
        super().__init__(**kwargs)
        assert layout in ("TNC", "NTC")
        self._hidden_size, self._num_layers, self._layout, self._dropout, self._dir = hidden_size, num_layers, layout, dropout, 2 if bidirectional else 1
        self._cells = []
        with self.name_scope():
            for layer in range(num_layers):
                for d in range(self._dir):
                    c = type(self)._cell(hidden_size, prefix="%s%d_" % ("lr"[d], layer), **cell_kw)
                    self._cells.append(c)
                    self.register_child(c)

    def state_info(self, batch_size=0):
        n = len(self._cells[0].state_info())
        return [{"shape": (self._num_layers * self._dir, batch_size, self._hidden_size), "__layout__": "LNC"}] * n

    def begin_state(self, batch_size=0, func=None, ctx=None, **kw):
        func = func or nd.zeros
        return [func(i["shape"], ctx=ctx) for i in self.state_info(batch_size)]

    def forward(self, inputs, states=None):
        x = inputs._t if self._layout == "TNC" else inputs._t.transpose(0, 1)
        T, N = x.shape[0], x.shape[1]
        ret_states = states is not None
        states = states or self.begin_state(N, ctx=inputs.context)
        states = [states] if isinstance(states, NDArray) else states
        n_state = len(states)
        finals = [[] for _ in range(n_state)]
        seq = x
        for layer in range(self._num_layers):
            outs_dir = []
            for d in range(self._dir):
                idx = layer * self._dir + d
                cell = self._cells[idx]
                st = [NDArray(s._t[idx]) for s in states]
                steps = range(T) if d == 0 else range(T - 1, -1, -1)
                outs = [None] * T
                for t in steps:
                    o, st = cell(NDArray(seq[t]), st)
                    outs[t] = o._t
                outs_dir.append(torch.stack(outs, 0))
                for k in range(n_state):
                    finals[k].append(st[k]._t)
            seq = torch.cat(outs_dir, dim=-1) if self._dir == 2 else outs_dir[0]
            if self._dropout and layer + 1 < self._num_layers:
                from ... import autograd
                seq = torch.nn.functional.dropout(seq, self._dropout, autograd.is_training())
        out = NDArray(seq if self._layout == "TNC" else seq.transpose(0, 1))
        if not ret_states:
            return out
        return out, [NDArray(torch.stack(f, 0)) for f in finals]


class RNN(_RNNLayer):
    _cell = RNNCell

    def __init__(self, hidden_size, num_layers=1, activation="relu", **kwargs):
        super().__init__(hidden_size, num_layers, activation=activation, **kwargs)


class LSTM(_RNNLayer):
    _cell = LSTMCell


class GRU(_RNNLayer):
    _cell = GRUCell
