"""``mx.gluon.contrib.data`` — IntervalSampler (parity: python/mxnet/gluon/contrib/data/sampler.py:25-70)."""
from __future__ import annotations

from ..data.sampler import Sampler

__all__ = ["IntervalSampler"]


class IntervalSampler(Sampler):
    """Indices ``0, k, 2k, …`` then (with ``rollover``) ``1, k+1, …`` and so on until every index in ``[0, length)`` was visited."""

    def __init__(self, length, interval, rollover=True):
        assert interval < length, "interval %d must be smaller than length %d" % (interval, length)
        self._length, self._interval, self._rollover = length, interval, rollover

    def __iter__(self):
        for start in range(self._interval if self._rollover else 1):
            yield from range(start, self._length, self._interval)

    def __len__(self):
        return self._length if self._rollover else (self._length + self._interval - 1) // self._interval
