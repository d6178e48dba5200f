"""Vision datasets: MNIST / FashionMNIST / CIFAR10.

Parity: ``python/mxnet/gluon/data/vision/datasets.py`` (MNIST reads ``train-images-idx3-ubyte.gz`` &c. from
``root``; items are ``(HxWx1 uint8 image, int32 label)``).  There is no network here, so when the idx / cifar
files are absent under ``root`` a deterministic **synthetic** dataset of the same shape/dtype/size is produced
(class-dependent blobs, so a model can actually learn it — used by tests and the demo scripts); set
``GEOMX_SYNTHETIC_SIZE`` to shrink it.  Parsing of real idx files is native (``_C.read_idx``) when available."""
from __future__ import annotations

import gzip
import os
import struct

import numpy as np
import torch

from ....base import getenv_int
from ....ndarray import NDArray
from ..dataset import Dataset

__all__ = ["MNIST", "FashionMNIST", "CIFAR10", "CIFAR100", "ImageRecordDataset", "ImageFolderDataset", "SyntheticImageDataset"]


def _read_idx(path):
    if not path.endswith(".gz"):
        from .... import runtime
        if runtime.available():                      # native parser (csrc/runtime/io.h): one read, no Python-level copies of the payload
            dims, raw = runtime.C().read_idx(path)
            return np.frombuffer(raw, dtype=np.uint8).reshape(dims)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    magic = struct.unpack(">I", buf[:4])[0]
    ndim = magic & 0xFF
    dims = struct.unpack(">%dI" % ndim, buf[4:4 + 4 * ndim])
    return np.frombuffer(buf, dtype=np.uint8, offset=4 + 4 * ndim).reshape(dims)


def _synthetic(n, shape, num_classes, seed, split=0):
    """Class-conditional images: fixed random prototype per class + noise (learnable, deterministic).  The prototypes depend on the dataset
    seed only, so the train and the test split of one dataset describe the same classes; labels / noise differ per split."""
    protos = np.random.RandomState(seed).randint(0, 200, size=(num_classes,) + shape).astype(np.float32)
    rng = np.random.RandomState(seed * 1000 + 17 + split)
    labels = (np.arange(n) * 7 + rng.randint(0, num_classes)) % num_classes
    noise = rng.randint(0, 56, size=(n,) + shape).astype(np.float32)
    data = np.clip(protos[labels] * 0.8 + noise, 0, 255).astype(np.uint8)
    return data, labels.astype(np.int32)


class _ImgDataset(Dataset):
    def __init__(self, data, label, transform=None):
        self._data, self._label, self._transform = data, label, transform

    def __getitem__(self, idx):
        x = NDArray(torch.from_numpy(self._data[idx]))
        y = int(self._label[idx])
        if self._transform is not None:
            return self._transform(x, y)
        return x, y

    def __len__(self):
        return len(self._label)


class SyntheticImageDataset(_ImgDataset):
    def __init__(self, n=1024, shape=(28, 28, 1), num_classes=10, seed=0, transform=None):
        d, l = _synthetic(n, shape, num_classes, seed)
        super().__init__(d, l, transform)


class MNIST(_ImgDataset):
    _files = {True: ("train-images-idx3-ubyte", "train-labels-idx1-ubyte"),
              False: ("t10k-images-idx3-ubyte", "t10k-labels-idx1-ubyte")}
    _sizes = {True: 60000, False: 10000}
    _seed = 42

    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "mnist"), train=True, transform=None):
        root = os.path.expanduser(root)
        data = label = None
        im, lb = self._files[train]
        for ext in ("", ".gz"):
            p_im, p_lb = os.path.join(root, im + ext), os.path.join(root, lb + ext)
            if os.path.exists(p_im) and os.path.exists(p_lb):
                data = _read_idx(p_im)[..., None].copy(); label = _read_idx(p_lb).astype(np.int32)
                break
        self.synthetic = data is None
        if data is None:
            n = getenv_int("GEOMX_SYNTHETIC_SIZE", 0) or self._sizes[train]
            if not train:
                n = max(1, n // 6)
            data, label = _synthetic(n, (28, 28, 1), 10, self._seed, 0 if train else 1)
        super().__init__(data, label, transform)


class FashionMNIST(MNIST):
    _seed = 43

    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "fashion-mnist"), train=True, transform=None):
        super().__init__(root, train, transform)


class CIFAR10(_ImgDataset):
    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "cifar10"), train=True, transform=None):
        root = os.path.expanduser(root)
        files = ["data_batch_%d.bin" % i for i in range(1, 6)] if train else ["test_batch.bin"]
        paths = [os.path.join(root, f) for f in files]
        if all(os.path.exists(p) for p in paths):
            raw = np.concatenate([np.fromfile(p, dtype=np.uint8).reshape(-1, 3072 + 1) for p in paths])
            label = raw[:, 0].astype(np.int32)
            data = raw[:, 1:].reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1).copy()
            self.synthetic = False
        else:
            n = getenv_int("GEOMX_SYNTHETIC_SIZE", 0) or (50000 if train else 10000)
            data, label = _synthetic(n, (32, 32, 3), 10, 44, 0 if train else 1)
            self.synthetic = True
        super().__init__(data, label, transform)


class CIFAR100(_ImgDataset):
    """CIFAR-100 from ``train.bin`` / ``test.bin`` under ``root`` (rows: coarse label, fine label, 3072 pixels); ``fine_label`` selects
    which of the two is returned.  Synthetic stand-in when the files are absent."""

    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "cifar100"), fine_label=False, train=True, transform=None):
        root = os.path.expanduser(root)
        path = os.path.join(root, "train.bin" if train else "test.bin")
        if os.path.exists(path):
            raw = np.fromfile(path, dtype=np.uint8).reshape(-1, 3072 + 2)
            label = raw[:, 1 if fine_label else 0].astype(np.int32)
            data = raw[:, 2:].reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1).copy()
            self.synthetic = False
        else:
            n = getenv_int("GEOMX_SYNTHETIC_SIZE", 0) or (50000 if train else 10000)
            data, label = _synthetic(n, (32, 32, 3), 100 if fine_label else 20, 45, 0 if train else 1)
            self.synthetic = True
        super().__init__(data, label, transform)


class ImageRecordDataset(Dataset):
    """Images + labels from an indexed RecordIO file packed by ``tools/im2rec.py`` (datasets.py ImageRecordDataset :230-270)."""

    def __init__(self, filename, flag=1, transform=None):
        from ..dataset import RecordFileDataset
        self._rec, self._flag, self._transform = RecordFileDataset(filename), flag, transform

    def __getitem__(self, idx):
        from .... import image, recordio
        header, img = recordio.unpack(self._rec[idx])
        x = image.imdecode(img, self._flag)
        y = header.label
        return self._transform(x, y) if self._transform is not None else (x, y)

    def __len__(self):
        return len(self._rec)


class ImageFolderDataset(Dataset):
    """``root/<class name>/<image>``: classes are the sorted sub-directory names (``synsets``), items ``(image HWC uint8, class index)``."""

    def __init__(self, root, flag=1, transform=None):
        self._root, self._flag, self._transform = os.path.expanduser(root), flag, transform
        self._exts = (".jpg", ".jpeg", ".png", ".bmp", ".ppm")
        self.synsets, self.items = [], []
        for folder in sorted(os.listdir(self._root)):
            path = os.path.join(self._root, folder)
            if not os.path.isdir(path):
                continue
            label = len(self.synsets)
            self.synsets.append(folder)
            for fn in sorted(os.listdir(path)):
                if os.path.splitext(fn)[1].lower() in self._exts:
                    self.items.append((os.path.join(path, fn), label))

    def __getitem__(self, idx):
        from .... import image
        x = image.imread(self.items[idx][0], self._flag)
        y = self.items[idx][1]
        return self._transform(x, y) if self._transform is not None else (x, y)

    def __len__(self):
        return len(self.items)
