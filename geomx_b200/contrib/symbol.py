"""``mx.contrib.symbol`` — the contrib operators as symbolic nodes (alias of ``mx.sym.contrib``, see the generic imperative-op bridge in
``geomx_b200/symbol.py``)."""
from __future__ import annotations


def __getattr__(name):
    from .. import symbol as _sym
    return getattr(_sym.contrib, name)
