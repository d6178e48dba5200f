"""``mx.contrib.io.DataLoaderIter`` — presents a ``gluon.data.DataLoader`` as a ``DataIter`` so symbolic ``Module.fit`` can consume it
(parity: python/mxnet/contrib/io.py:25-95).  The last, smaller batch is padded up to ``batch_size`` and ``pad`` reports how many rows are
filler."""
from __future__ import annotations

import torch

from ..io import DataBatch, DataDesc, DataIter
from ..ndarray import NDArray

__all__ = ["DataLoaderIter"]


class DataLoaderIter(DataIter):
    def __init__(self, loader, data_name="data", label_name="softmax_label", dtype="float32"):
        super().__init__()
        self._loader, self._iter, self._dtype = loader, iter(loader), dtype
        data, label = next(self._iter)
        self.batch_size = data.shape[0]
        self.provide_data = [DataDesc(data_name, tuple(data.shape), dtype)]
        self.provide_label = [DataDesc(label_name, tuple(label.shape), dtype)]
        self._cur, self._first = None, (data, label)

    def reset(self):
        self._iter, self._first = iter(self._loader), None

    def iter_next(self):
        if self._first is not None:
            self._cur, self._first = self._first, None
            return True
        try:
            self._cur = next(self._iter)
            return True
        except StopIteration:
            self._cur = None
            return False

    def next(self):
        if not self.iter_next():
            raise StopIteration
        return DataBatch(self.getdata(), self.getlabel(), pad=self.getpad(), index=None, provide_data=self.provide_data, provide_label=self.provide_label)

    def _padded(self, arr):
        t = arr._t
        if t.shape[0] < self.batch_size:
            t = torch.cat([t, t.new_zeros((self.batch_size - t.shape[0],) + tuple(t.shape[1:]))])
        return [NDArray(t).astype(self._dtype)]

    def getdata(self):
        return self._padded(self._cur[0])

    def getlabel(self):
        return self._padded(self._cur[1])

    def getpad(self):
        return self.batch_size - self._cur[0].shape[0]

    def getindex(self):
        return None
