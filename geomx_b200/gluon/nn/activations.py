"""``mx.gluon.nn.activations`` — the activation blocks under the reference's module path (``python/mxnet/gluon/nn/activations.py``)."""
from .basic_layers import ELU, SELU, Activation, LeakyReLU, PReLU, Swish  # noqa: F401

__all__ = ["Activation", "LeakyReLU", "PReLU", "ELU", "SELU", "Swish"]
