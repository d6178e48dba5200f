"""Multi-layer (bi)directional recurrent layers over the cells' parameters (the reference dispatches to cuDNN's fused RNN).  Per layer and
direction the input projections of all time steps are ONE GEMM (``T*N`` rows — large enough for the tcgen05 path on CUDA); the time loop
only carries the hidden-to-hidden GEMM and the gate arithmetic."""
from __future__ import annotations

import torch

from ... import ndarray as nd
from ...ndarray import NDArray
from ...ops import functional as OF
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
                # the input projections of ALL time steps are one GEMM (T*N rows); the recurrence only carries the h2h GEMM
                cell._deferred_infer(NDArray(seq[0]))
                ctx = inputs.context
                wi, wh, bi, bh = (p.data(ctx)._t for p in (cell.i2h_weight, cell.h2h_weight, cell.i2h_bias, cell.h2h_bias))
                xi_all = OF.dense(seq.reshape(T * N, -1), wi, bi, None, False).reshape(T, N, -1)
                h = st[0]._t
                c = st[1]._t if len(st) > 1 else None
                kind = type(self)._cell.__name__
                for t in steps:
                    gh = OF.dense(h, wh, bh, None, False)
                    if kind == "LSTMCell":
                        i, f, g, o = (xi_all[t] + gh).chunk(4, dim=-1)
                        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                        h = torch.sigmoid(o) * torch.tanh(c)
                    elif kind == "GRUCell":
                        xr, xz, xn = xi_all[t].chunk(3, dim=-1); hr, hz, hn = gh.chunk(3, dim=-1)
                        r = torch.sigmoid(xr + hr); z = torch.sigmoid(xz + hz)
                        h = (1 - z) * torch.tanh(xn + r * hn) + z * h
                    else:
                        h = OF.activation(xi_all[t] + gh, cell._activation)
                    outs[t] = h
                st = [NDArray(h)] + ([NDArray(c)] if c is not None else [])
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
