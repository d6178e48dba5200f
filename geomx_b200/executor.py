"""``mx.executor`` — the bound-graph executor (parity: python/mxnet/executor.py ``Executor``: forward / backward / arg_dict / grad_dict /
aux_dict / outputs / copy_params_from).  The implementation lives next to the graph it runs (``geomx_b200/symbol.py``)."""
from .symbol import Executor  # noqa: F401

__all__ = ["Executor"]
