"""Checkpoints of models built from ``mx.rnn`` cells (parity: python/mxnet/rnn/rnn.py).  Files always hold the UNPACKED per-gate weights, so a
model trained with ``FusedRNNCell`` loads into its ``unfuse()``-d form and back."""
from ..model import load_checkpoint, save_checkpoint
from .rnn_cell import BaseRNNCell

__all__ = ["save_rnn_checkpoint", "load_rnn_checkpoint", "do_rnn_checkpoint"]


def _cells(cells):
    return [cells] if isinstance(cells, BaseRNNCell) else list(cells)


def save_rnn_checkpoint(cells, prefix, epoch, symbol, arg_params, aux_params):
    """``save_checkpoint`` after unpacking the weights of every cell in ``cells``."""
    for c in _cells(cells):
        arg_params = c.unpack_weights(arg_params)
    save_checkpoint(prefix, epoch, symbol, arg_params, aux_params)


def load_rnn_checkpoint(cells, prefix, epoch):
    """``load_checkpoint`` followed by packing the weights into the layout the given cells use.  Returns (symbol, arg_params, aux_params)."""
    symbol, arg, aux = load_checkpoint(prefix, epoch)
    for c in _cells(cells):
        arg = c.pack_weights(arg)
    return symbol, arg, aux


def do_rnn_checkpoint(cells, prefix, period=1):
    """Epoch-end callback for ``Module.fit`` that checkpoints every ``period`` epochs with unpacked weights."""
    period = int(max(1, period))

    def _callback(iter_no, symbol, arg, aux):
        if (iter_no + 1) % period == 0:
            save_rnn_checkpoint(cells, prefix, iter_no + 1, symbol, arg, aux)
    return _callback
