"""``mx.name`` — automatic symbol naming (parity: python/mxnet/name.py: ``NameManager`` hands out ``<hint><counter>`` names, ``Prefix`` prepends
a prefix inside a ``with`` block)."""
from __future__ import annotations

import threading

__all__ = ["NameManager", "Prefix"]


class NameManager:
    _tls = threading.local()

    def __init__(self):
        self._counter, self._old = {}, None

    def get(self, name, hint):
        if name:
            return name
        i = self._counter.get(hint, 0)
        self._counter[hint] = i + 1
        return "%s%d" % (hint, i)

    def __enter__(self):
        self._old = getattr(NameManager._tls, "current", None)
        NameManager._tls.current = self
        return self

    def __exit__(self, *exc):
        NameManager._tls.current = self._old

    @staticmethod
    def current():
        cur = getattr(NameManager._tls, "current", None)
        if cur is None:
            cur = NameManager._tls.current = NameManager()
        return cur


class Prefix(NameManager):
    def __init__(self, prefix):
        super().__init__()
        self._prefix = prefix

    def get(self, name, hint):
        return self._prefix + super().get(name, hint)
