"""Data iterators (``mx.io``).  Parity: ``python/mxnet/io.py`` (DataDesc, DataBatch, DataIter, NDArrayIter,
ResizeIter, PrefetchingIter) and the C++ ``MNISTIter`` (``src/io/iter_mnist.cc:80-260``: idx files → batches
(N,1,28,28) scaled to [0,1], optional flat/shuffle/partition ``part_index/num_parts``)."""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from .ndarray import NDArray

__all__ = ["DataDesc", "DataBatch", "DataIter", "NDArrayIter", "MNISTIter", "ResizeIter", "CSVIter", "LibSVMIter", "ImageRecordIter", "PrefetchingIter"]

class DataDesc(namedtuple("DataDesc", ["name", "shape", "dtype", "layout"])):
    """Name, shape, dtype and layout of one input (``python/mxnet/io/io.py`` DataDesc)."""
    __slots__ = ()

    def __new__(cls, name, shape, dtype="float32", layout="NCHW"):
        return super().__new__(cls, name, tuple(shape), dtype, layout)

    def __repr__(self):
        return "DataDesc[%s,%s,%s,%s]" % (self.name, self.shape, self.dtype, self.layout)

    @staticmethod
    def get_batch_axis(layout):
        """Index of the batch dimension ``N`` in ``layout`` (0 when the layout is unknown)."""
        return 0 if layout is None else layout.find("N")

    @staticmethod
    def get_list(shapes, types):
        """``[(name, shape), ...]`` (+ ``[(name, dtype), ...]``) -> list of DataDesc."""
        if types is not None:
            tdict = dict(types)
            return [DataDesc(n, s, tdict[n]) for n, s in shapes]
        return [DataDesc(n, s) for n, s in shapes]


class DataBatch:
    def __init__(self, data, label=None, pad=None, index=None, provide_data=None, provide_label=None, bucket_key=None):
        self.data, self.label, self.pad, self.index = data, label, pad, index
        self.provide_data, self.provide_label, self.bucket_key = provide_data, provide_label, bucket_key


class DataIter:
    def __init__(self, batch_size=0):
        self.batch_size = batch_size

    def __iter__(self):
        return self

    def reset(self):
        pass

    def next(self):
        """Default ``next`` for iterators written against the low-level protocol (``iter_next`` + ``getdata`` / ``getlabel`` / ``getpad`` /
        ``getindex``, io.py:180-260); high-level iterators override ``next`` directly."""
        if type(self).iter_next is not DataIter.iter_next and self.iter_next():
            return DataBatch(data=self.getdata(), label=self.getlabel(), pad=self.getpad(), index=self.getindex())
        raise StopIteration

    def __next__(self):
        return self.next()

    def iter_next(self):
        return False

    def getdata(self):
        return None

    def getlabel(self):
        return None

    def getindex(self):
        return None

    def getpad(self):
        return 0


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


class CSVIter(NDArrayIter):
    """``mx.io.CSVIter(data_csv=, data_shape=, label_csv=, label_shape=, batch_size=)`` — src/io/iter_csv.cc; parsed natively (text_io.h)."""

    def __init__(self, data_csv, data_shape, label_csv=None, label_shape=(1,), batch_size=1, round_batch=True, **kw):
        data = _parse_csv(data_csv).reshape((-1,) + tuple(data_shape))
        label = None
        if label_csv is not None:
            label = _parse_csv(label_csv).reshape((-1,) + tuple(label_shape))
            if tuple(label_shape) == (1,):
                label = label.reshape(-1)
        else:
            label = np.zeros((data.shape[0],), dtype=np.float32)
        super().__init__(data, label, batch_size=batch_size, shuffle=False, last_batch_handle="pad" if round_batch else "discard", **kw)


def _parse_csv(path):
    from . import runtime
    if runtime.available():
        return np.asarray(runtime.C().parse_csv(str(path)), dtype=np.float32)
    return np.loadtxt(path, delimiter=",", dtype=np.float32, ndmin=2)


class LibSVMIter(NDArrayIter):
    """``mx.io.LibSVMIter(data_libsvm=, data_shape=(D,), batch_size=)`` — src/io/iter_libsvm.cc.  The native parser returns CSR; batches are
    served dense (row_sparse/CSR consumers convert with ``tostype``)."""

    def __init__(self, data_libsvm, data_shape, batch_size=1, label_libsvm=None, round_batch=True, **kw):
        from . import runtime
        D = int(data_shape[0] if isinstance(data_shape, (tuple, list)) else data_shape)
        if runtime.available():
            labels, values, indices, indptr, _ = runtime.C().parse_libsvm(str(data_libsvm))
        else:  # pragma: no cover - pure python fallback
            labels, values, indices, indptr = [], [], [], [0]
            for line in open(data_libsvm):
                parts = line.split("#")[0].split()
                if not parts:
                    continue
                labels.append(float(parts[0]))
                for kvp in parts[1:]:
                    i, v = kvp.split(":"); indices.append(int(i)); values.append(float(v))
                indptr.append(len(indices))
        n = len(labels)
        dense = np.zeros((n, D), dtype=np.float32)
        indptr = np.asarray(indptr); indices = np.asarray(indices); values = np.asarray(values, dtype=np.float32)
        rows = np.repeat(np.arange(n), np.diff(indptr))
        dense[rows, indices] = values
        super().__init__(dense, np.asarray(labels, dtype=np.float32), batch_size=batch_size, shuffle=False,
                         last_batch_handle="pad" if round_batch else "discard", **kw)


class PrefetchingIter(DataIter):
    """Background-thread prefetch of another iterator (``python/mxnet/io.py`` PrefetchingIter; C++ src/io/iter_prefetcher.h): the producer
    runs ``depth`` batches ahead so decoding / parsing overlaps the training step."""

    def __init__(self, base, depth=2):
        import queue
        super().__init__(getattr(base, "batch_size", 0))
        self._base, self._depth, self._q, self._thread = base, depth, queue.Queue(maxsize=depth), None
        self._start()

    provide_data = property(lambda self: self._base.provide_data)
    provide_label = property(lambda self: self._base.provide_label)

    def _start(self):
        import threading

        def run(q, base):
            try:
                for b in base:
                    q.put(b)
            finally:
                q.put(None)
        self._thread = threading.Thread(target=run, args=(self._q, self._base), daemon=True)
        self._thread.start()

    def reset(self):
        if self._thread is not None:
            while self._thread.is_alive():                    # drain so that the producer can finish
                try:
                    self._q.get(timeout=0.05)
                except Exception:
                    pass
        import queue
        self._q = queue.Queue(maxsize=self._depth)
        self._base.reset()
        self._start()

    def next(self):
        b = self._q.get()
        if b is None:
            raise StopIteration
        return b


class ImageRecordIter(DataIter):
    """``mx.io.ImageRecordIter(path_imgrec=, data_shape=(C,H,W), batch_size=, shuffle=, rand_crop=, rand_mirror=, mean_r/g/b=, scale=)`` over a
    RecordIO image file (src/io/iter_image_recordio_2.cc).  Records are decoded with Pillow on a prefetch thread; ``path_imgidx`` enables
    shuffling without scanning."""

    def __init__(self, path_imgrec, data_shape, batch_size, path_imgidx=None, shuffle=False, rand_crop=False, rand_mirror=False,
                 mean_r=0.0, mean_g=0.0, mean_b=0.0, scale=1.0, resize=0, label_width=1, seed=0, prefetch=2, data_name="data",
                 label_name="softmax_label", **kw):
        from . import recordio
        super().__init__(batch_size)
        self._shape, self._shuffle, self._crop, self._mirror = tuple(data_shape), shuffle, rand_crop, rand_mirror
        self._mean = np.array([mean_r, mean_g, mean_b], dtype=np.float32)[: self._shape[0]].reshape(-1, 1, 1)
        self._scale, self._resize, self._lw = scale, resize, label_width
        self._rng = np.random.RandomState(seed)
        self.data_name, self.label_name = data_name, label_name
        if path_imgidx:
            self._rec = recordio.MXIndexedRecordIO(path_imgidx, path_imgrec, "r")
            self._keys = list(self._rec.keys)
        else:
            self._rec = recordio.MXRecordIO(path_imgrec, "r")
            self._keys = None
        self._records = None
        if self._keys is None:                                 # sequential file: load the raw records once (shuffling needs random access)
            self._records = []
            while True:
                r = self._rec.read()
                if r is None:
                    break
                self._records.append(r)
        self.reset()

    @property
    def provide_data(self):
        return [DataDesc(self.data_name, (self.batch_size,) + self._shape)]

    @property
    def provide_label(self):
        return [DataDesc(self.label_name, (self.batch_size,) if self._lw == 1 else (self.batch_size, self._lw))]

    def reset(self):
        n = len(self._keys) if self._keys is not None else len(self._records)
        self._order = self._rng.permutation(n) if self._shuffle else np.arange(n)
        self._cursor = 0

    def _decode(self, raw):
        from PIL import Image
        from . import recordio
        header, img = recordio.unpack_img(raw, iscolor=1 if self._shape[0] == 3 else 0)
        C, H, W = self._shape
        im = Image.fromarray(img)
        if self._resize:
            w, h = im.size
            s = self._resize / float(min(w, h))
            im = im.resize((max(1, int(round(w * s))), max(1, int(round(h * s)))))
        w, h = im.size
        if w < W or h < H:
            im = im.resize((max(w, W), max(h, H))); w, h = im.size
        x0 = self._rng.randint(0, w - W + 1) if self._crop else (w - W) // 2
        y0 = self._rng.randint(0, h - H + 1) if self._crop else (h - H) // 2
        arr = np.asarray(im.crop((x0, y0, x0 + W, y0 + H)), dtype=np.float32)
        arr = arr[None, :, :] if arr.ndim == 2 else arr.transpose(2, 0, 1)
        if self._mirror and self._rng.rand() < 0.5:
            arr = arr[:, :, ::-1]
        return (arr - self._mean) * self._scale, header.label

    def next(self):
        n = len(self._order)
        if self._cursor >= n:
            raise StopIteration
        idx = self._order[self._cursor:self._cursor + self.batch_size]
        pad = self.batch_size - len(idx)
        if pad:
            idx = np.concatenate([idx, self._order[:pad]])
        self._cursor += self.batch_size
        xs, ys = [], []
        for i in idx:
            raw = self._rec.read_idx(self._keys[i]) if self._keys is not None else self._records[i]
            x, y = self._decode(raw)
            xs.append(x); ys.append(y)
        data = NDArray(torch.from_numpy(np.ascontiguousarray(np.stack(xs))))
        label = NDArray(torch.from_numpy(np.asarray(ys, dtype=np.float32)))
        return DataBatch([data], [label], pad=pad)


import abc as _abc  # noqa: E402


class MXDataIter(DataIter, metaclass=_abc.ABCMeta):
    """In the reference this wraps an iterator implemented in C++ (``MXDataIterCreateIter``).  The native iterators of this framework
    (``MNISTIter``, ``CSVIter``, ``LibSVMIter``, ``ImageRecordIter`` over ``csrc/runtime/{io,text_io}.h``) are ordinary ``DataIter`` classes, so
    this name is kept as the (virtual) base they are recognised by: ``isinstance(it, MXDataIter)`` is true for exactly those."""

    @_abc.abstractmethod
    def next(self):
        raise NotImplementedError


for _cls in (MNISTIter, CSVIter, LibSVMIter, ImageRecordIter):
    MXDataIter.register(_cls)
del _cls


# the reference's package layout (python/mxnet/io/{io,utils}.py) as importable paths
def _register_paths():
    from ._alias import submodule
    g = globals()
    names = [k for k, v in g.items() if isinstance(v, type) and (issubclass(v, DataIter) or k in ("DataBatch", "DataDesc"))]
    submodule(__name__, "io", {k: g[k] for k in names})
    submodule(__name__, "utils", {k: g[k] for k in ("_init_data", "_has_instance", "_getdata_by_idx") if k in g})


_register_paths()
