"""``mx.AttrScope`` — attributes attached to every symbol created inside the scope (parity: python/mxnet/attribute.py; used for
``ctx_group`` / ``lr_mult`` / ``wd_mult`` annotations)."""
from __future__ import annotations

import threading

__all__ = ["AttrScope"]


class AttrScope:
    _tls = threading.local()

    def __init__(self, **kwargs):
        for v in kwargs.values():
            if not isinstance(v, str):
                raise ValueError("Attributes need to be string")
        self._attr, self._old = kwargs, None

    def get(self, attr=None):
        out = dict(self._attr)
        out.update(attr or {})
        return out

    def __enter__(self):
        self._old = getattr(AttrScope._tls, "current", None)
        merged = AttrScope(**(self._old.get(self._attr) if self._old is not None else self._attr))
        AttrScope._tls.current = merged
        return self

    def __exit__(self, *exc):
        AttrScope._tls.current = self._old

    @staticmethod
    def current():
        cur = getattr(AttrScope._tls, "current", None)
        if cur is None:
            cur = AttrScope._tls.current = AttrScope()
        return cur
