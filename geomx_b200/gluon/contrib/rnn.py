"""``mx.gluon.contrib.rnn`` — VariationalDropoutCell, LSTMPCell and the convolutional recurrent cells.

Parity: ``python/mxnet/gluon/contrib/rnn/rnn_cell.py`` (VariationalDropoutCell :27-180, LSTMPCell :183-320) and ``conv_rnn_cell.py``
(Conv{1,2,3}D{RNN,LSTM,GRU}Cell: gates computed by an input-to-hidden and a hidden-to-hidden convolution, gate order as in the dense cells)."""
from __future__ import annotations

import torch
import torch.nn.functional as TF

from ... import autograd
from ... import ndarray as nd
from ...ndarray import NDArray
from ...ops import functional as OF
from ..block import HybridBlock
from ..rnn.rnn_cell import ModifierCell, RecurrentCell

__all__ = ["VariationalDropoutCell", "LSTMPCell", "Conv1DRNNCell", "Conv2DRNNCell", "Conv3DRNNCell", "Conv1DLSTMCell", "Conv2DLSTMCell",
           "Conv3DLSTMCell", "Conv1DGRUCell", "Conv2DGRUCell", "Conv3DGRUCell"]


class VariationalDropoutCell(ModifierCell):
    """Dropout with ONE mask per sequence (Gal & Ghahramani) on inputs / states / outputs; call ``reset()`` between sequences."""

    def __init__(self, base_cell, drop_inputs=0.0, drop_states=0.0, drop_outputs=0.0, **kwargs):
        super().__init__(base_cell, **kwargs)
        self._di, self._ds, self._do = drop_inputs, drop_states, drop_outputs
        self.reset()

    def reset(self):
        self._mi = self._ms = self._mo = None

    @staticmethod
    def _mask(p, like):
        return (torch.rand_like(like) >= p).to(like.dtype) / (1.0 - p)

    def forward(self, inputs, states):
        if not autograd.is_training():
            return self.base_cell(inputs, states)
        if self._di > 0:
            if self._mi is None:
                self._mi = self._mask(self._di, inputs._t)
            inputs = NDArray(inputs._t * self._mi)
        if self._ds > 0:
            if self._ms is None:
                self._ms = self._mask(self._ds, states[0]._t)
            states = [NDArray(states[0]._t * self._ms)] + list(states[1:])
        out, nxt = self.base_cell(inputs, states)
        if self._do > 0:
            if self._mo is None:
                self._mo = self._mask(self._do, out._t)
            out = NDArray(out._t * self._mo)
        return out, nxt

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        return super().unroll(length, inputs, begin_state, layout, merge_outputs)


class LSTMPCell(RecurrentCell):
    """LSTM with a projection of the hidden state (Sak et al. 2014): states are ``(r [N, P], c [N, H])``, ``r = W_hr · h``."""
    _gates = 4

    def __init__(self, hidden_size, projection_size, input_size=0, i2h_weight_initializer=None, h2h_weight_initializer=None,
                 h2r_weight_initializer=None, i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", **kwargs):
        HybridBlock.__init__(self, **kwargs)
        from ..nn.basic_layers import _bias_init
        self._hidden_size, self._projection_size, self._input_size = hidden_size, projection_size, input_size
        with self.name_scope():
            self.i2h_weight = self.params.get("i2h_weight", shape=(4 * hidden_size, input_size), init=i2h_weight_initializer, allow_deferred_init=True)
            self.h2h_weight = self.params.get("h2h_weight", shape=(4 * hidden_size, projection_size), init=h2h_weight_initializer, allow_deferred_init=True)
            self.h2r_weight = self.params.get("h2r_weight", shape=(projection_size, hidden_size), init=h2r_weight_initializer, allow_deferred_init=True)
            self.i2h_bias = self.params.get("i2h_bias", shape=(4 * hidden_size,), init=_bias_init(i2h_bias_initializer), allow_deferred_init=True)
            self.h2h_bias = self.params.get("h2h_bias", shape=(4 * hidden_size,), init=_bias_init(h2h_bias_initializer), allow_deferred_init=True)

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size, self._projection_size), "__layout__": "NC"}, {"shape": (batch_size, self._hidden_size), "__layout__": "NC"}]

    def _deferred_infer(self, x):
        if self.i2h_weight.shape[1] == 0 or self.i2h_weight._data is None:
            self.i2h_weight.shape = (4 * self._hidden_size, x.shape[-1])
            for prm in (self.i2h_weight, self.h2h_weight, self.h2r_weight, self.i2h_bias, self.h2h_bias):
                prm._finish_deferred_init() if hasattr(prm, "_finish_deferred_init") else None

    def forward(self, inputs, states):
        self._deferred_infer(inputs)
        c = inputs.context
        wi, wh, wr, bi, bh = (p.data(c)._t for p in (self.i2h_weight, self.h2h_weight, self.h2r_weight, self.i2h_bias, self.h2h_bias))
        i, f, g, o = (TF.linear(inputs._t, wi, bi) + TF.linear(states[0]._t, wh, bh)).chunk(4, dim=-1)
        cc = torch.sigmoid(f) * states[1]._t + torch.sigmoid(i) * torch.tanh(g)
        r = TF.linear(torch.sigmoid(o) * torch.tanh(cc), wr)
        return NDArray(r), [NDArray(r), NDArray(cc)]


class _ConvRNNBase(RecurrentCell):
    _gates, _nd = 1, 2

    def __init__(self, input_shape, hidden_channels, i2h_kernel, h2h_kernel, i2h_pad=0, i2h_dilate=1, h2h_dilate=1, i2h_weight_initializer=None,
                 h2h_weight_initializer=None, i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", conv_layout=None, activation="tanh", **kwargs):
        HybridBlock.__init__(self, **kwargs)
        from ..nn.basic_layers import _bias_init
        n = self._nd
        tup = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v,) * n   # noqa: E731
        self._input_shape, self._hc, self._act = tuple(input_shape), hidden_channels, activation
        self._ik, self._hk, self._ip, self._id, self._hd = tup(i2h_kernel), tup(h2h_kernel), tup(i2h_pad), tup(i2h_dilate), tup(h2h_dilate)
        assert all(k % 2 == 1 for k in self._hk), "h2h_kernel must be odd so the state keeps its shape"
        self._hp = tuple(d * (k - 1) // 2 for k, d in zip(self._hk, self._hd))
        spatial = tuple((s + 2 * p - d * (k - 1) - 1) + 1 for s, p, d, k in zip(self._input_shape[1:], self._ip, self._id, self._ik))
        self._state_shape = (hidden_channels,) + spatial
        G = self._gates * hidden_channels
        with self.name_scope():
            self.i2h_weight = self.params.get("i2h_weight", shape=(G, self._input_shape[0]) + self._ik, init=i2h_weight_initializer)
            self.h2h_weight = self.params.get("h2h_weight", shape=(G, hidden_channels) + self._hk, init=h2h_weight_initializer)
            self.i2h_bias = self.params.get("i2h_bias", shape=(G,), init=_bias_init(i2h_bias_initializer))
            self.h2h_bias = self.params.get("h2h_bias", shape=(G,), init=_bias_init(h2h_bias_initializer))

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size,) + self._state_shape, "__layout__": "NC" + "DHW"[-self._nd:]}] * (2 if self._gates == 4 else 1)

    def _deferred_infer(self, x):
        pass

    def _convs(self, x, h, wi, wh, bi, bh):
        conv = (TF.conv1d, TF.conv2d, TF.conv3d)[self._nd - 1]
        return conv(x._t, wi._t, bi._t, 1, self._ip, self._id), conv(h._t, wh._t, bh._t, 1, self._hp, self._hd)

    def _step(self, x, states, wi, wh, bi, bh):
        a, b = self._convs(x, states[0], wi, wh, bi, bh)
        if self._gates == 1:
            h = NDArray(OF.activation(a + b, self._act))
            return h, [h]
        if self._gates == 4:
            i, f, g, o = (a + b).chunk(4, dim=1)
            c = torch.sigmoid(f) * states[1]._t + torch.sigmoid(i) * OF.activation(g, self._act)
            h = torch.sigmoid(o) * OF.activation(c, self._act)
            return NDArray(h), [NDArray(h), NDArray(c)]
        ai, bi_ = a.chunk(3, dim=1), b.chunk(3, dim=1)
        r = torch.sigmoid(ai[0] + bi_[0]); z = torch.sigmoid(ai[1] + bi_[1])
        n = OF.activation(ai[2] + r * bi_[2], self._act)
        h = (1 - z) * n + z * states[0]._t
        return NDArray(h), [NDArray(h)]


def _make(name, gates, nd_):
    return type(name, (_ConvRNNBase,), {"_gates": gates, "_nd": nd_, "__doc__": "%s: input ``(C, *spatial)``, see ``_ConvRNNBase``." % name})


Conv1DRNNCell, Conv2DRNNCell, Conv3DRNNCell = (_make("Conv%dDRNNCell" % d, 1, d) for d in (1, 2, 3))
Conv1DLSTMCell, Conv2DLSTMCell, Conv3DLSTMCell = (_make("Conv%dDLSTMCell" % d, 4, d) for d in (1, 2, 3))
Conv1DGRUCell, Conv2DGRUCell, Conv3DGRUCell = (_make("Conv%dDGRUCell" % d, 3, d) for d in (1, 2, 3))
