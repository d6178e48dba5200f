"""Datasets.  Parity: ``python/mxnet/gluon/data/dataset.py`` (Dataset.transform / transform_first, SimpleDataset,
ArrayDataset, _LazyTransformDataset)."""
from __future__ import annotations

__all__ = ["Dataset", "SimpleDataset", "ArrayDataset", "RecordFileDataset"]


class Dataset:
    def __getitem__(self, idx):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def transform(self, fn, lazy=True):
        trans = _LazyTransformDataset(self, fn)
        return trans if lazy else SimpleDataset([i for i in trans])

    def transform_first(self, fn, lazy=True):
        def base_fn(x, *args):
            return (fn(x),) + args if args else fn(x)
        return self.transform(base_fn, lazy)


class SimpleDataset(Dataset):
    def __init__(self, data):
        self._data = data

    def __len__(self):
        return len(self._data)

    def __getitem__(self, idx):
        return self._data[idx]


class _LazyTransformDataset(Dataset):
    def __init__(self, data, fn):
        self._data, self._fn = data, fn

    def __len__(self):
        return len(self._data)

    def __getitem__(self, idx):
        item = self._data[idx]
        return self._fn(*item) if isinstance(item, tuple) else self._fn(item)


class ArrayDataset(Dataset):
    def __init__(self, *args):
        assert len(args) > 0, "Needs at least 1 arrays"
        self._length = len(args[0])
        self._data = []
        for i, d in enumerate(args):
            assert len(d) == self._length, "All arrays must have the same length; array[0] has length %d while array[%d] has %d." % (self._length, i + 1, len(d))
            self._data.append(d)

    def __getitem__(self, idx):
        return self._data[0][idx] if len(self._data) == 1 else tuple(d[idx] for d in self._data)

    def __len__(self):
        return self._length



class RecordFileDataset(Dataset):
    """Raw records of an indexed RecordIO file (``.rec`` + ``.idx``) (dataset.py RecordFileDataset :170-200)."""

    def __init__(self, filename):
        import os
        from ... import recordio
        self.idx_file = os.path.splitext(filename)[0] + ".idx"
        self.filename = filename
        self._record = recordio.MXIndexedRecordIO(self.idx_file, self.filename, "r")

    def __getitem__(self, idx):
        return self._record.read_idx(self._record.keys[idx])

    def __len__(self):
        return len(self._record.keys)
