"""``mx.image`` — image decoding, resizing / cropping helpers, augmenters and ``ImageIter`` (parity: python/mxnet/image/image.py: imread :45,
imdecode :85, scale_down, resize_short :197, fixed_crop :258, random_crop :291, center_crop :331, color_normalize :380, Augmenter classes
:480-900, CreateAugmenter :904, ImageIter :1010-1300).  Decoding uses Pillow (the reference uses OpenCV); arrays are HWC uint8/float NDArrays
like the reference's."""
from __future__ import annotations

import os
import random as _pyrandom

import numpy as np
import torch

from . import io as _io
from . import ndarray as nd
from .ndarray import NDArray

__all__ = ["imread", "imdecode", "imresize", "resize_short", "fixed_crop", "center_crop", "random_crop", "color_normalize", "Augmenter",
           "ResizeAug", "ForceResizeAug", "RandomCropAug", "CenterCropAug", "HorizontalFlipAug", "CastAug", "ColorNormalizeAug", "BrightnessJitterAug",
           "CreateAugmenter", "ImageIter"]


def _pil():
    from PIL import Image
    return Image


def imdecode(buf, flag=1, to_rgb=1):
    import io
    img = _pil().open(io.BytesIO(bytes(buf)))
    img = img.convert("RGB" if flag else "L")
    arr = np.asarray(img)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    if flag and not to_rgb:
        arr = arr[:, :, ::-1]
    return nd.array(np.ascontiguousarray(arr), dtype="uint8")


def imread(filename, flag=1, to_rgb=1):
    with open(filename, "rb") as f:
        return imdecode(f.read(), flag, to_rgb)


def imresize(src, w, h, interp=1):
    arr = src.asnumpy()
    squeeze = arr.shape[2] == 1
    img = _pil().fromarray(arr[:, :, 0] if squeeze else arr.astype(np.uint8))
    out = np.asarray(img.resize((int(w), int(h)), _pil().BILINEAR if interp == 1 else _pil().NEAREST))
    return nd.array(out[:, :, None] if squeeze else out, dtype=str(arr.dtype))


def resize_short(src, size, interp=1):
    h, w = src.shape[:2]
    if h > w:
        return imresize(src, size, int(h * size / w), interp)
    return imresize(src, int(w * size / h), size, interp)


def fixed_crop(src, x0, y0, w, h, size=None, interp=1):
    out = NDArray(src._t[y0:y0 + h, x0:x0 + w].clone())
    if size is not None and (w, h) != tuple(size):
        out = imresize(out, size[0], size[1], interp)
    return out


def center_crop(src, size, interp=1):
    h, w = src.shape[:2]
    nw, nh = min(w, size[0]), min(h, size[1])
    x0, y0 = (w - nw) // 2, (h - nh) // 2
    return fixed_crop(src, x0, y0, nw, nh, size, interp), (x0, y0, nw, nh)


def random_crop(src, size, interp=1):
    h, w = src.shape[:2]
    nw, nh = min(w, size[0]), min(h, size[1])
    x0, y0 = _pyrandom.randint(0, w - nw), _pyrandom.randint(0, h - nh)
    return fixed_crop(src, x0, y0, nw, nh, size, interp), (x0, y0, nw, nh)


def color_normalize(src, mean, std=None):
    out = src.astype("float32") - (mean if isinstance(mean, NDArray) else nd.array(np.asarray(mean, dtype=np.float32)))
    if std is not None:
        out = out / (std if isinstance(std, NDArray) else nd.array(np.asarray(std, dtype=np.float32)))
    return out


class Augmenter:
    def __init__(self, **kwargs):
        self._kwargs = kwargs

    def dumps(self):
        import json
        return json.dumps([self.__class__.__name__.lower(), self._kwargs])

    def __call__(self, src):
        raise NotImplementedError


class ResizeAug(Augmenter):
    def __init__(self, size, interp=1): super().__init__(size=size, interp=interp); self.size, self.interp = size, interp
    def __call__(self, src): return resize_short(src, self.size, self.interp)


class ForceResizeAug(Augmenter):
    def __init__(self, size, interp=1): super().__init__(size=size, interp=interp); self.size, self.interp = size, interp
    def __call__(self, src): return imresize(src, self.size[0], self.size[1], self.interp)


class RandomCropAug(Augmenter):
    def __init__(self, size, interp=1): super().__init__(size=size, interp=interp); self.size, self.interp = size, interp
    def __call__(self, src): return random_crop(src, self.size, self.interp)[0]


class CenterCropAug(Augmenter):
    def __init__(self, size, interp=1): super().__init__(size=size, interp=interp); self.size, self.interp = size, interp
    def __call__(self, src): return center_crop(src, self.size, self.interp)[0]


class HorizontalFlipAug(Augmenter):
    def __init__(self, p): super().__init__(p=p); self.p = p
    def __call__(self, src): return NDArray(src._t.flip(1)) if _pyrandom.random() < self.p else src


class CastAug(Augmenter):
    def __init__(self, typ="float32"): super().__init__(type=typ); self.typ = typ
    def __call__(self, src): return src.astype(self.typ)


class ColorNormalizeAug(Augmenter):
    def __init__(self, mean, std): super().__init__(mean=list(np.ravel(mean)), std=None if std is None else list(np.ravel(std))); self.mean, self.std = mean, std
    def __call__(self, src): return color_normalize(src, self.mean, self.std)


class BrightnessJitterAug(Augmenter):
    def __init__(self, brightness): super().__init__(brightness=brightness); self.brightness = brightness
    def __call__(self, src): return src.astype("float32") * (1.0 + _pyrandom.uniform(-self.brightness, self.brightness))


_GRAY = (0.299, 0.587, 0.114)


def _gray(t):
    return (t[..., 0] * _GRAY[0] + t[..., 1] * _GRAY[1] + t[..., 2] * _GRAY[2]).unsqueeze(-1)


class ContrastJitterAug(Augmenter):
    def __init__(self, contrast): super().__init__(contrast=contrast); self.contrast = contrast

    def __call__(self, src):
        t = src._t.float(); alpha = 1.0 + _pyrandom.uniform(-self.contrast, self.contrast)
        return NDArray(t * alpha + _gray(t).mean() * (1.0 - alpha))


class SaturationJitterAug(Augmenter):
    def __init__(self, saturation): super().__init__(saturation=saturation); self.saturation = saturation

    def __call__(self, src):
        t = src._t.float(); alpha = 1.0 + _pyrandom.uniform(-self.saturation, self.saturation)
        return NDArray(t * alpha + _gray(t) * (1.0 - alpha))


class HueJitterAug(Augmenter):
    """Rotate the hue by a random angle in YIQ space (image.py HueJitterAug :760-800)."""

    def __init__(self, hue):
        super().__init__(hue=hue); self.hue = hue
        self.tyiq = np.array([[0.299, 0.587, 0.114], [0.596, -0.274, -0.321], [0.211, -0.523, 0.311]])
        self.ityiq = np.array([[1.0, 0.956, 0.621], [1.0, -0.272, -0.647], [1.0, -1.107, 1.705]])

    def __call__(self, src):
        alpha = _pyrandom.uniform(-self.hue, self.hue)
        u, w = np.cos(alpha * np.pi), np.sin(alpha * np.pi)
        bt = np.array([[1.0, 0.0, 0.0], [0.0, u, -w], [0.0, w, u]])
        m = torch.as_tensor(np.dot(np.dot(self.ityiq, bt), self.tyiq).T, dtype=torch.float32)
        return NDArray(src._t.float() @ m.to(src._t.device))


class ColorJitterAug(Augmenter):
    """Brightness, contrast and saturation jitter applied in random order."""

    def __init__(self, brightness, contrast, saturation):
        super().__init__(brightness=brightness, contrast=contrast, saturation=saturation)
        self.ts = [a for a, v in ((BrightnessJitterAug, brightness), (ContrastJitterAug, contrast), (SaturationJitterAug, saturation)) if v > 0]
        self.ts = [a(v) for a, v in zip(self.ts, [v for v in (brightness, contrast, saturation) if v > 0])]

    def __call__(self, src):
        order = list(self.ts); _pyrandom.shuffle(order)
        for t in order:
            src = t(src)
        return src


class LightingAug(Augmenter):
    """AlexNet-style PCA lighting noise: adds ``eigvec @ (alpha * eigval)`` with ``alpha ~ N(0, alphastd)``."""

    def __init__(self, alphastd, eigval, eigvec):
        super().__init__(alphastd=alphastd, eigval=list(np.ravel(eigval)), eigvec=[list(r) for r in np.asarray(eigvec)])
        self.alphastd, self.eigval, self.eigvec = alphastd, np.asarray(eigval, dtype=np.float32), np.asarray(eigvec, dtype=np.float32)

    def __call__(self, src):
        alpha = np.random.normal(0, self.alphastd, size=(3,)).astype(np.float32)
        rgb = np.dot(self.eigvec * alpha, self.eigval)
        return NDArray(src._t.float() + torch.as_tensor(rgb, device=src._t.device))


class RandomGrayAug(Augmenter):
    def __init__(self, p): super().__init__(p=p); self.p = p

    def __call__(self, src):
        if _pyrandom.random() < self.p:
            return NDArray(_gray(src._t.float()).expand(-1, -1, 3).contiguous())
        return src


class RandomOrderAug(Augmenter):
    def __init__(self, ts): super().__init__(); self.ts = list(ts)

    def dumps(self):
        return [self.__class__.__name__.lower(), [t.dumps() for t in self.ts]]

    def __call__(self, src):
        order = list(self.ts); _pyrandom.shuffle(order)
        for t in order:
            src = t(src)
        return src


class SequentialAug(Augmenter):
    def __init__(self, ts): super().__init__(); self.ts = list(ts)

    def dumps(self):
        return [self.__class__.__name__.lower(), [t.dumps() for t in self.ts]]

    def __call__(self, src):
        for t in self.ts:
            src = t(src)
        return src


def scale_down(src_size, size):
    """Shrink ``size`` (w, h) proportionally so that it fits into ``src_size`` (w, h)."""
    w, h = size; sw, sh = src_size
    if sh < h:
        w, h = float(w * sh) / h, sh
    if sw < w:
        w, h = sw, float(h * sw) / w
    return int(w), int(h)


def random_size_crop(src, size, area, ratio, interp=2, **kwargs):
    """Random crop with ``area`` fraction (float or (min, max)) and aspect ``ratio`` range, resized to ``size``; falls back to a centre
    crop after 10 failed draws.  Returns ``(image, (x0, y0, w, h))``."""
    import math
    h, w = src.shape[0], src.shape[1]
    src_area = h * w
    if "min_area" in kwargs:
        area = kwargs.pop("min_area")
    if isinstance(area, (int, float)):
        area = (area, 1.0)
    for _ in range(10):
        target = _pyrandom.uniform(area[0], area[1]) * src_area
        log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
        ar = math.exp(_pyrandom.uniform(*log_ratio))
        nw, nh = int(round(math.sqrt(target * ar))), int(round(math.sqrt(target / ar)))
        if nw <= w and nh <= h:
            x0, y0 = _pyrandom.randint(0, w - nw), _pyrandom.randint(0, h - nh)
            return fixed_crop(src, x0, y0, nw, nh, size, interp), (x0, y0, nw, nh)
    return center_crop(src, size, interp)


class RandomSizedCropAug(Augmenter):
    def __init__(self, size, area, ratio, interp=2, **kwargs):
        super().__init__(size=size, area=area, ratio=ratio, interp=interp)
        self.size, self.area, self.ratio, self.interp = size, kwargs.pop("min_area", area), ratio, interp

    def __call__(self, src):
        return random_size_crop(src, self.size, self.area, self.ratio, self.interp)[0]


def CreateAugmenter(data_shape, resize=0, rand_crop=False, rand_resize=False, rand_mirror=False, mean=None, std=None, brightness=0, contrast=0,
                    saturation=0, hue=0, pca_noise=0, rand_gray=0, inter_method=1, **kwargs):
    augs = []
    if resize > 0:
        augs.append(ResizeAug(resize, inter_method))
    crop = (data_shape[2], data_shape[1])
    if rand_resize:
        assert rand_crop
        augs.append(RandomSizedCropAug(crop, 0.08, (3.0 / 4.0, 4.0 / 3.0), inter_method))
    else:
        augs.append(RandomCropAug(crop, inter_method) if rand_crop else CenterCropAug(crop, inter_method))
    if rand_mirror:
        augs.append(HorizontalFlipAug(0.5))
    augs.append(CastAug())
    if brightness or contrast or saturation:
        augs.append(ColorJitterAug(brightness, contrast, saturation))
    if hue:
        augs.append(HueJitterAug(hue))
    if pca_noise > 0:
        augs.append(LightingAug(pca_noise, np.array([55.46, 4.794, 1.148]),
                                np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]])))
    if rand_gray > 0:
        augs.append(RandomGrayAug(rand_gray))
    if mean is True:
        mean = np.array([123.68, 116.28, 103.53])
    if std is True:
        std = np.array([58.395, 57.12, 57.375])
    if mean is not None:
        augs.append(ColorNormalizeAug(mean, std))
    return augs


class ImageIter(_io.DataIter):
    """Images from a RecordIO file (``path_imgrec`` [+ ``path_imgidx``]) or an image list (``imglist`` / ``path_imglist`` + ``path_root``)
    through a list of augmenters; yields NCHW float batches."""

    def __init__(self, batch_size, data_shape, label_width=1, path_imgrec=None, path_imglist=None, path_root="", path_imgidx=None, shuffle=False,
                 aug_list=None, imglist=None, data_name="data", label_name="softmax_label", **kwargs):
        super().__init__(batch_size)
        from . import recordio
        self.data_shape, self.label_width, self.shuffle = tuple(data_shape), label_width, shuffle
        self.data_name, self.label_name = data_name, label_name
        self.auglist = aug_list if aug_list is not None else CreateAugmenter(data_shape, **kwargs)
        self._rec, self._items = None, []
        if path_imgrec:
            if path_imgidx:
                self._rec = recordio.MXIndexedRecordIO(path_imgidx, path_imgrec, "r")
                self._items = list(self._rec.keys)
            else:
                # no index file: find the records in the memory-mapped file (native scan) and fetch payloads lazily by offset — the
                # payloads themselves are never all resident
                self._reader = recordio.RecordReader(path_imgrec)
                self._items = [("@", off) for off in self._reader.offsets]
        else:
            entries = imglist
            if entries is None:
                entries = []
                for line in open(path_imglist):
                    p = line.strip().split("\t")
                    entries.append([float(x) for x in p[1:-1]] + [p[-1]])
            for e in entries:
                self._items.append((np.asarray(e[:-1], dtype=np.float32), os.path.join(path_root, e[-1])))
        self.reset()

    @property
    def provide_data(self): return [_io.DataDesc(self.data_name, (self.batch_size,) + self.data_shape)]
    @property
    def provide_label(self): return [_io.DataDesc(self.label_name, (self.batch_size,) if self.label_width == 1 else (self.batch_size, self.label_width))]

    def reset(self):
        self._order = list(range(len(self._items)))
        if self.shuffle:
            _pyrandom.shuffle(self._order)
        self._cur = 0

    # ---- the per-sample pipeline as overridable steps (python/mxnet/image/image.py ImageIter: next_sample / imdecode / check_valid_image /
    # augmentation_transform / postprocess_data), so that sub-classes such as ImageDetIter change one stage only
    def hard_reset(self):
        """Reset ignoring any roll-over state (this iterator keeps none)."""
        self.reset()

    def _fetch(self, i):
        """``(label, raw)`` of item ``i``: ``raw`` is the encoded image (bytes) or, for image lists, the file path."""
        from . import recordio
        it = self._items[i]
        if isinstance(it, tuple) and len(it) == 2 and isinstance(it[0], str) and it[0] == "@":
            it = self._reader.read(it[1])
        if self._rec is not None or isinstance(it, bytes):
            header, img = recordio.unpack(self._rec.read_idx(it) if self._rec is not None else it)
            return header.label, img
        return (it[0][0] if self.label_width == 1 else it[0]), it[1]

    def next_sample(self):
        """``(label, raw)`` of the next sample in iteration order (raises StopIteration at the end of the epoch)."""
        if self._cur >= len(self._order):
            raise StopIteration
        self._cur += 1
        return self._fetch(self._order[self._cur - 1])

    def read_image(self, fname):
        with open(fname, "rb") as f:
            return f.read()

    def imdecode(self, s):
        """Decode raw bytes (or read + decode a path) into an HWC uint8 NDArray."""
        if isinstance(s, str):
            s = self.read_image(s)
        return imdecode(s, flag=1 if self.data_shape[0] == 3 else 0)

    def check_valid_image(self, data):
        if len(data[0].shape) == 0:
            raise RuntimeError("Data shape is wrong")

    def check_data_shape(self, data_shape):
        if not len(data_shape) == 3:
            raise ValueError("data_shape should have length 3, with dimensions CxHxW")
        if not data_shape[0] in (1, 3):
            raise ValueError("This iterator expects inputs to have 1 or 3 channels.")

    def augmentation_transform(self, data):
        for aug in self.auglist:
            data = aug(data)
        return data

    def postprocess_data(self, datum):
        """HWC -> CHW float tensor of the batch."""
        return datum._t.permute(2, 0, 1).float()

    def _sample(self, i):
        label, raw = self._fetch(i)
        arr = self.imdecode(raw)
        self.check_valid_image([arr])
        return self.postprocess_data(self.augmentation_transform(arr)), label

    def next(self):
        if self._cur >= len(self._order):
            raise StopIteration
        idx = self._order[self._cur:self._cur + self.batch_size]
        pad = self.batch_size - len(idx)
        idx = idx + self._order[:pad]
        self._cur += self.batch_size
        xs, ys = zip(*[self._sample(i) for i in idx])
        return _io.DataBatch([NDArray(torch.stack(xs))], [nd.array(np.asarray(ys, dtype=np.float32))], pad=pad)


from .image_detection import *  # noqa: E402,F401,F403  (mx.image.ImageDetIter, Det*Aug, CreateDetAugmenter)


# the reference's package layout (python/mxnet/image/{image,detection}.py) as importable paths
def _register_paths():
    import sys
    from ._alias import submodule
    from . import image_detection as _det
    g = globals()
    own = {k: v for k, v in g.items() if not k.startswith("_") and getattr(v, "__module__", None) == __name__}
    submodule(__name__, "image", own)
    sys.modules[__name__ + ".detection"] = _det
    g["detection"] = _det


_register_paths()
