"""``mx.symbol_doc`` — see :mod:`ndarray_doc` (reference: ``python/mxnet/symbol_doc.py``)."""
from .ndarray_doc import SymbolDoc, _build_doc  # noqa: F401

__all__ = ["SymbolDoc", "_build_doc"]
