"""``mx.gluon.rnn`` — recurrent cells and layers (parity: python/mxnet/gluon/rnn/{rnn_cell,rnn_layer}.py: RNNCell / LSTMCell / GRUCell with
``i2h``/``h2h`` weights in MXNet's gate order, ``unroll``, and the fused-layer classes RNN / LSTM / GRU with ``layout`` TNC|NTC, multi-layer,
bidirectional, ``begin_state``)."""
from .rnn_cell import (BidirectionalCell, DropoutCell, GRUCell, HybridRecurrentCell, HybridSequentialRNNCell, LSTMCell, ModifierCell, RecurrentCell, ResidualCell, RNNCell,  # noqa: F401
                       SequentialRNNCell, ZoneoutCell)
from .rnn_layer import GRU, LSTM, RNN  # noqa: F401
