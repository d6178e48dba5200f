"""Data iterators (``mx.io``).  Parity: ``python/mxnet/io.py`` (DataDesc, DataBatch, DataIter, NDArrayIter,
ResizeIter, PrefetchingIter) and the C++ ``MNISTIter`` (``src/io/iter_mnist.cc:80-260``: idx files → batches
(N,1,28,28) scaled to [0,1], optional flat/shuffle/partition ``part_index/num_parts``)."""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from .ndarray import NDArray, array

__all__ = ["DataDesc", "DataBatch", "DataIter", "NDArrayIter", "MNISTIter", "ResizeIter"]

DataDesc = namedtuple("DataDesc", ["name", "shape", "dtype", "layout"])
DataDesc.__new__.__defaults__ = ("float32", "NCHW")


class DataBatch:
    def __init__(self, data, label=None, pad=None, index=None, provide_data=None, provide_label=None):
        self.data, self.label, self.pad, self.index = data, label, pad, index
        self.provide_data, self.provide_label = provide_data, provide_label


class DataIter:
    def __init__(self, batch_size=0):
        self.batch_size = batch_size

    def __iter__(self):
        return self

    def reset(self):
        pass

    def next(self):
        raise StopIteration

    def __next__(self):
        return self.next()


class NDArrayIter(DataIter):
    def __init__(self, data, label=None, batch_size=1, shuffle=False, last_batch_handle="pad",
                 data_name="data", label_name="softmax_label"):
        super().__init__(batch_size)
        self._data = data._t if isinstance(data, NDArray) else torch.as_tensor(np.asarray(data))
        self._label = None if label is None else (label._t if isinstance(label, NDArray) else torch.as_tensor(np.asarray(label)))
        self._shuffle, self._lbh = shuffle, last_batch_handle
        self._n = self._data.shape[0]
        self.data_name, self.label_name = data_name, label_name
        self.reset()

    @property
    def provide_data(self):
        return [DataDesc(self.data_name, (self.batch_size,) + tuple(self._data.shape[1:]))]

    @property
    def provide_label(self):
        return [] if self._label is None else [DataDesc(self.label_name, (self.batch_size,) + tuple(self._label.shape[1:]))]

    def reset(self):
        self._cursor = 0
        self._perm = torch.randperm(self._n) if self._shuffle else torch.arange(self._n)

    def next(self):
        if self._cursor >= self._n:
            raise StopIteration
        idx = self._perm[self._cursor:self._cursor + self.batch_size]
        pad = self.batch_size - idx.numel()
        if pad:
            if self._lbh == "discard":
                raise StopIteration
            if self._lbh == "pad":
                idx = torch.cat([idx, self._perm[:pad]])
        self._cursor += self.batch_size
        d = [NDArray(self._data[idx])]
        l = None if self._label is None else [NDArray(self._label[idx])]
        return DataBatch(d, l, pad=pad if self._lbh == "pad" else 0)


class MNISTIter(NDArrayIter):
    def __init__(self, image="./train-images-idx3-ubyte", label="./train-labels-idx1-ubyte", batch_size=128,
                 shuffle=True, flat=False, seed=0, silent=False, num_parts=1, part_index=0, **kw):
        import os
        from .gluon.data.vision.datasets import _read_idx, _synthetic
        if os.path.exists(image) and os.path.exists(label):
            img = _read_idx(image).astype(np.float32) / 255.0; lab = _read_idx(label).astype(np.float32)
        else:
            d, l = _synthetic(60000 if "train" in image else 10000, (28, 28), 10, 42)
            img = d.astype(np.float32) / 255.0; lab = l.astype(np.float32)
        n = img.shape[0] // num_parts
        img, lab = img[part_index * n:(part_index + 1) * n], lab[part_index * n:(part_index + 1) * n]
        img = img.reshape(n, -1) if flat else img.reshape(n, 1, 28, 28)
        if shuffle:
            torch.manual_seed(seed)
        super().__init__(img, lab, batch_size, shuffle, "discard")


class ResizeIter(DataIter):
    def __init__(self, data_iter, size, reset_internal=True):
        super().__init__(data_iter.batch_size)
        self.data_iter, self.size, self.reset_internal, self.cur = data_iter, size, reset_internal, 0

    def reset(self):
        self.cur = 0
        if self.reset_internal:
            self.data_iter.reset()

    def next(self):
        if self.cur == self.size:
            raise StopIteration
        try:
            b = self.data_iter.next()
        except StopIteration:
            self.data_iter.reset(); b = self.data_iter.next()
        self.cur += 1
        return b
