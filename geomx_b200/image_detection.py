"""Object-detection data pipeline: augmenters that move the boxes with the pixels, and ``ImageDetIter``.

Parity: ``python/mxnet/image/detection.py`` (DetAugmenter, DetBorrowAug, DetRandomSelectAug, DetHorizontalFlipAug, DetRandomCropAug,
DetRandomPadAug, CreateMultiRandCropAugmenter, CreateDetAugmenter, ImageDetIter).  Labels are ``[n, w]`` float arrays, one object per
row: ``(class id, xmin, ymin, xmax, ymax, extras…)`` with coordinates normalised to [0, 1].  In a record / image list the raw label is
``[header width A, object width B, (A-2 extra header values), obj0 (B values), obj1, …]``; batches carry ``[batch, max objects, B]``
padded with -1."""
from __future__ import annotations

import json
import random as _pyrandom

import numpy as np
import torch

from . import io as _io
from . import ndarray as nd
from .image import (Augmenter, BrightnessJitterAug, CastAug, ColorJitterAug, ColorNormalizeAug, ForceResizeAug, HueJitterAug, ImageIter,
                    LightingAug, RandomGrayAug, ResizeAug, fixed_crop)
from .ndarray import NDArray

__all__ = ["DetAugmenter", "DetBorrowAug", "DetRandomSelectAug", "DetHorizontalFlipAug", "DetRandomCropAug", "DetRandomPadAug",
           "CreateMultiRandCropAugmenter", "CreateDetAugmenter", "ImageDetIter"]


class DetAugmenter:
    """``aug(image, label) -> (image, label)``"""

    def __init__(self, **kwargs):
        self._kwargs = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in kwargs.items()}

    def dumps(self):
        return json.dumps([self.__class__.__name__.lower(), self._kwargs])

    def __call__(self, src, label):
        raise NotImplementedError("Must override implementation.")


class DetBorrowAug(DetAugmenter):
    """Use a plain image augmenter that does not move pixels relative to the boxes (colour, cast, force-resize)."""

    def __init__(self, augmenter):
        if not isinstance(augmenter, Augmenter):
            raise TypeError("Borrowing from invalid Augmenter")
        super().__init__(augmenter=augmenter.dumps())
        self.augmenter = augmenter

    def dumps(self):
        return [self.__class__.__name__.lower(), self.augmenter.dumps()]

    def __call__(self, src, label):
        return self.augmenter(src), label


class DetRandomSelectAug(DetAugmenter):
    """Apply ONE randomly chosen augmenter of the list, or none with probability ``skip_prob``."""

    def __init__(self, aug_list, skip_prob=0):
        super().__init__(skip_prob=skip_prob)
        if not isinstance(aug_list, (list, tuple)):
            aug_list = [aug_list]
        for a in aug_list:
            if not isinstance(a, DetAugmenter):
                raise ValueError("Allow DetAugmenter in list only")
        if not aug_list:
            skip_prob = 1
        self.aug_list, self.skip_prob = list(aug_list), skip_prob

    def dumps(self):
        return [self.__class__.__name__.lower(), [a.dumps() for a in self.aug_list]]

    def __call__(self, src, label):
        if _pyrandom.random() < self.skip_prob:
            return src, label
        return _pyrandom.choice(self.aug_list)(src, label)


class DetHorizontalFlipAug(DetAugmenter):
    def __init__(self, p):
        super().__init__(p=p); self.p = p

    def __call__(self, src, label):
        if _pyrandom.random() < self.p:
            src = NDArray(src._t.flip(1))
            label = label.copy()
            xmin = 1.0 - label[:, 3]
            label[:, 3] = 1.0 - label[:, 1]
            label[:, 1] = xmin
        return src, label


def _areas(boxes):
    return np.maximum(0, boxes[:, 2] - boxes[:, 0]) * np.maximum(0, boxes[:, 3] - boxes[:, 1])


def _intersect(boxes, x1, y1, x2, y2):
    out = boxes.copy()
    out[:, 0] = np.maximum(boxes[:, 0], x1); out[:, 1] = np.maximum(boxes[:, 1], y1)
    out[:, 2] = np.minimum(boxes[:, 2], x2); out[:, 3] = np.minimum(boxes[:, 3], y2)
    bad = (out[:, 0] >= out[:, 2]) | (out[:, 1] >= out[:, 3])
    out[bad] = 0
    return out


class DetRandomCropAug(DetAugmenter):
    """Random crop constrained by object coverage: the crop must cover at least ``min_object_covered`` of some object; objects whose
    remaining visible fraction falls below ``min_eject_coverage`` are dropped, the others are clipped and re-normalised."""

    def __init__(self, min_object_covered=0.1, aspect_ratio_range=(0.75, 1.33), area_range=(0.05, 1.0), min_eject_coverage=0.3, max_attempts=50):
        if not isinstance(aspect_ratio_range, (tuple, list)):
            aspect_ratio_range = (aspect_ratio_range, aspect_ratio_range)
        if not isinstance(area_range, (tuple, list)):
            area_range = (area_range, area_range)
        super().__init__(min_object_covered=min_object_covered, aspect_ratio_range=aspect_ratio_range, area_range=area_range,
                         min_eject_coverage=min_eject_coverage, max_attempts=max_attempts)
        self.min_object_covered, self.min_eject_coverage, self.max_attempts = min_object_covered, min_eject_coverage, max_attempts
        self.aspect_ratio_range, self.area_range = tuple(aspect_ratio_range), tuple(area_range)
        self.enabled = area_range[1] > 0 and area_range[0] <= area_range[1] and aspect_ratio_range[0] > 0 and aspect_ratio_range[0] <= aspect_ratio_range[1]

    def __call__(self, src, label):
        crop = self._random_crop_proposal(label, src.shape[0], src.shape[1])
        if crop:
            x, y, w, h, label = crop
            src = fixed_crop(src, x, y, w, h, None)
        return src, label

    def _update_labels(self, label, crop_box, height, width):
        xmin, ymin, w, h = (float(crop_box[0]) / width, float(crop_box[1]) / height, float(crop_box[2]) / width, float(crop_box[3]) / height)
        inter = _intersect(label[:, 1:5], xmin, ymin, xmin + w, ymin + h)
        cov = _areas(inter) / np.maximum(_areas(label[:, 1:5]), 1e-12)
        valid = (cov > self.min_eject_coverage) & (_areas(inter) * width * height > 2)
        if not valid.any():
            return None
        out = label[valid].copy()
        b = inter[valid]
        out[:, 1] = (b[:, 0] - xmin) / w; out[:, 2] = (b[:, 1] - ymin) / h
        out[:, 3] = (b[:, 2] - xmin) / w; out[:, 4] = (b[:, 3] - ymin) / h
        return out

    def _random_crop_proposal(self, label, height, width):
        if not self.enabled or height <= 0 or width <= 0:
            return ()
        min_area, max_area = self.area_range[0] * height * width, self.area_range[1] * height * width
        for _ in range(self.max_attempts):
            ratio = _pyrandom.uniform(*self.aspect_ratio_range)
            if ratio <= 0:
                continue
            h = int(round(np.sqrt(min_area / ratio))); max_h = int(round(np.sqrt(max_area / ratio)))
            if round(max_h * ratio) > width:
                max_h = int((width + 0.4999999) / ratio)
            max_h = min(max_h, height)
            h = min(h, max_h)
            if h < max_h:
                h = _pyrandom.randint(h, max_h)
            w = int(round(h * ratio))
            if w <= 0 or h <= 0 or w > width or h > height or w * h < min_area * 0.98 or w * h > max_area * 1.02:
                continue
            y, x = _pyrandom.randint(0, max(0, height - h)), _pyrandom.randint(0, max(0, width - w))
            boxes = label[:, 1:5]
            inter = _intersect(boxes, x / width, y / height, (x + w) / width, (y + h) / height)
            cov = _areas(inter) / np.maximum(_areas(boxes), 1e-12)
            cov = cov[_areas(boxes) * width * height > 2]
            if cov.size == 0 or np.amin(cov) <= self.min_object_covered and np.amax(cov) <= self.min_object_covered:
                continue
            new_label = self._update_labels(label, (x, y, w, h), height, width)
            if new_label is not None:
                return x, y, w, h, new_label
        return ()


class DetRandomPadAug(DetAugmenter):
    """Place the image at a random position of a larger canvas filled with ``pad_val``; boxes shrink accordingly."""

    def __init__(self, aspect_ratio_range=(0.75, 1.33), area_range=(1.0, 3.0), max_attempts=50, pad_val=(128, 128, 128)):
        if not isinstance(pad_val, (list, tuple)):
            pad_val = (pad_val,)
        if not isinstance(aspect_ratio_range, (tuple, list)):
            aspect_ratio_range = (aspect_ratio_range, aspect_ratio_range)
        if not isinstance(area_range, (tuple, list)):
            area_range = (area_range, area_range)
        super().__init__(aspect_ratio_range=aspect_ratio_range, area_range=area_range, max_attempts=max_attempts, pad_val=pad_val)
        self.pad_val, self.aspect_ratio_range, self.area_range, self.max_attempts = pad_val, tuple(aspect_ratio_range), tuple(area_range), max_attempts
        self.enabled = area_range[1] > 1.0 and area_range[0] <= area_range[1] and aspect_ratio_range[0] > 0 and aspect_ratio_range[0] <= aspect_ratio_range[1]

    def __call__(self, src, label):
        height, width = src.shape[0], src.shape[1]
        pad = self._random_pad_proposal(label, height, width)
        if pad:
            x, y, w, h, label = pad
            t = src._t
            canvas = torch.empty((h, w, t.shape[2]), dtype=t.dtype, device=t.device)
            vals = list(self.pad_val) * t.shape[2] if len(self.pad_val) == 1 else list(self.pad_val)
            for c in range(t.shape[2]):
                canvas[:, :, c] = vals[c]
            canvas[y:y + height, x:x + width] = t
            src = NDArray(canvas)
        return src, label

    def _random_pad_proposal(self, label, height, width):
        if not self.enabled or height <= 0 or width <= 0:
            return ()
        min_area, max_area = self.area_range[0] * height * width, self.area_range[1] * height * width
        for _ in range(self.max_attempts):
            ratio = _pyrandom.uniform(*self.aspect_ratio_range)
            if ratio <= 0:
                continue
            h = int(round(np.sqrt(min_area / ratio))); max_h = int(round(np.sqrt(max_area / ratio)))
            if round(h * ratio) < width:
                h = int((width + 0.499999) / ratio)
            h = max(h, height)
            if h > max_h:
                h = max_h
            if h < max_h:
                h = _pyrandom.randint(h, max_h)
            w = int(round(h * ratio))
            if (h - height) < 2 or (w - width) < 2:
                continue
            y, x = _pyrandom.randint(0, max(0, h - height)), _pyrandom.randint(0, max(0, w - width))
            out = label.copy()
            out[:, (1, 3)] = (out[:, (1, 3)] * width + x) / w
            out[:, (2, 4)] = (out[:, (2, 4)] * height + y) / h
            return x, y, w, h, out
        return ()


def CreateMultiRandCropAugmenter(min_object_covered=0.1, aspect_ratio_range=(0.75, 1.33), area_range=(0.05, 1.0), min_eject_coverage=0.3,
                                 max_attempts=50, skip_prob=0):
    """Several crop samplers (each argument may be a list, one entry per sampler) of which one is picked at random per image."""
    def align(params):
        out, num = [], 1
        for p in params:
            if not isinstance(p, list):
                p = [p]
            out.append(p); num = max(num, len(p))
        for k, p in enumerate(out):
            if len(p) != num:
                assert len(p) == 1
                out[k] = p * num
        return out
    aligned = align([min_object_covered, aspect_ratio_range, area_range, min_eject_coverage, max_attempts])
    augs = [DetRandomCropAug(min_object_covered=moc, aspect_ratio_range=arr, area_range=ar, min_eject_coverage=mec, max_attempts=ma)
            for moc, arr, ar, mec, ma in zip(*aligned)]
    return DetRandomSelectAug(augs, skip_prob=skip_prob)


def CreateDetAugmenter(data_shape, resize=0, rand_crop=0, rand_pad=0, rand_gray=0, rand_mirror=False, mean=None, std=None, brightness=0, contrast=0,
                       saturation=0, pca_noise=0, hue=0, inter_method=2, min_object_covered=0.1, aspect_ratio_range=(0.75, 1.33),
                       area_range=(0.05, 3.0), min_eject_coverage=0.3, max_attempts=50, pad_val=(127, 127, 127)):
    """The standard SSD-style chain: resize → random crop (prob ``rand_crop``) → mirror → random pad (prob ``rand_pad``) → force resize to
    ``data_shape`` → cast → colour jitter / hue / PCA noise / gray → normalise."""
    auglist = []
    if resize > 0:
        auglist.append(DetBorrowAug(ResizeAug(resize, inter_method)))
    if rand_crop > 0:
        auglist.append(CreateMultiRandCropAugmenter(min_object_covered, aspect_ratio_range, (area_range[0], min(1.0, area_range[1])), min_eject_coverage,
                                                    max_attempts, skip_prob=(1 - rand_crop)))
    if rand_mirror > 0:
        auglist.append(DetHorizontalFlipAug(0.5))
    if rand_pad > 0:
        auglist.append(DetRandomSelectAug([DetRandomPadAug(aspect_ratio_range, (1.0, area_range[1]), max_attempts, pad_val)], 1 - rand_pad))
    auglist.append(DetBorrowAug(ForceResizeAug((data_shape[2], data_shape[1]), inter_method)))
    auglist.append(DetBorrowAug(CastAug()))
    if brightness or contrast or saturation:
        auglist.append(DetBorrowAug(ColorJitterAug(brightness, contrast, saturation)))
    if hue:
        auglist.append(DetBorrowAug(HueJitterAug(hue)))
    if pca_noise > 0:
        auglist.append(DetBorrowAug(LightingAug(pca_noise, np.array([55.46, 4.794, 1.148]),
                                                np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]]))))
    if rand_gray > 0:
        auglist.append(DetBorrowAug(RandomGrayAug(rand_gray)))
    if mean is True:
        mean = np.array([123.68, 116.28, 103.53])
    if std is True:
        std = np.array([58.395, 57.12, 57.375])
    if mean is not None:
        auglist.append(DetBorrowAug(ColorNormalizeAug(mean, std)))
    return auglist


class ImageDetIter(ImageIter):
    """Detection batches: data ``[B, C, H, W]``, label ``[B, max objects, object width]`` padded with -1.  Sources as in ``ImageIter``."""

    def __init__(self, batch_size, data_shape, path_imgrec=None, path_imglist=None, path_root="", path_imgidx=None, shuffle=False, aug_list=None,
                 imglist=None, data_name="data", label_name="label", last_batch_handle="pad", **kwargs):
        super().__init__(batch_size=batch_size, data_shape=data_shape, label_width=-1, path_imgrec=path_imgrec, path_imglist=path_imglist, path_root=path_root,
                         path_imgidx=path_imgidx, shuffle=shuffle, aug_list=[], imglist=imglist, data_name=data_name, label_name=label_name)
        self.auglist = CreateDetAugmenter(data_shape, **kwargs) if aug_list is None else aug_list
        self.label_shape = self._estimate_label_shape()

    @property
    def provide_label(self):
        return [_io.DataDesc(self.label_name, (self.batch_size,) + self.label_shape)]

    @staticmethod
    def _parse_label(label):
        """Raw ``[A, B, header…, objects…]`` → ``[n, B]``; rows with an invalid box are dropped."""
        raw = np.asarray(label, dtype=np.float32).ravel()
        if raw.size < 7:
            raise RuntimeError("Label shape is invalid: " + str(raw.shape))
        header_width, obj_width = int(raw[0]), int(raw[1])
        if (raw.size - header_width) % obj_width != 0:
            raise RuntimeError("Label shape %s inconsistent with annotation width %d." % (str(raw.shape), obj_width))
        out = raw[header_width:].reshape(-1, obj_width)
        valid = (out[:, 3] > out[:, 1]) & (out[:, 4] > out[:, 2])
        if not valid.any():
            raise RuntimeError("Encounter sample with no valid label.")
        return out[valid]

    def _raw(self, i):
        from . import recordio
        from .image import imread
        it = self._items[i]
        if isinstance(it, tuple) and len(it) == 2 and isinstance(it[0], str) and it[0] == "@":       # lazily read record of an un-indexed .rec
            it = self._reader.read(it[1])
        if self._rec is not None or isinstance(it, bytes):
            header, img = recordio.unpack_img(self._rec.read_idx(it) if self._rec is not None else it, iscolor=1 if self.data_shape[0] == 3 else 0)
            return header.label, nd.array(img if img.ndim == 3 else img[:, :, None], dtype="uint8")
        return it[0], imread(it[1], flag=1 if self.data_shape[0] == 3 else 0)

    def _estimate_label_shape(self):
        max_count, width = 0, 5
        for i in range(len(self._items)):
            lab = self._parse_label(self._raw(i)[0])
            max_count, width = max(max_count, lab.shape[0]), lab.shape[1]
        return (max_count, width)

    def reshape(self, data_shape=None, label_shape=None):
        if data_shape is not None:
            self.data_shape = tuple(data_shape)
        if label_shape is not None:
            self.label_shape = tuple(label_shape)

    def sync_label_shape(self, it, verbose=False):
        """Make a train and a validation iterator agree on the (larger) label shape."""
        assert isinstance(it, ImageDetIter), "Synchronize with invalid iterator."
        shape = (max(self.label_shape[0], it.label_shape[0]), max(self.label_shape[1], it.label_shape[1]))
        self.reshape(None, shape); it.reshape(None, shape)
        return it

    def augmentation_transform(self, data, label):
        for aug in self.auglist:
            data, label = aug(data, label)
        return data, label

    def _sample(self, i):
        raw_label, arr = self._raw(i)
        arr, lab = self.augmentation_transform(arr, self._parse_label(raw_label))
        out = np.full(self.label_shape, -1.0, dtype=np.float32)
        n = min(lab.shape[0], self.label_shape[0])
        out[:n, :lab.shape[1]] = lab[:n]
        return arr._t.permute(2, 0, 1).float(), out

    def draw_next(self, color=None, thickness=2, mean=None, std=None, clip=True, waitKey=None, window_name="draw_next", id2labels=None):
        """Generator of HWC uint8 images with the ground-truth boxes burnt in (no GUI: the reference shows them with OpenCV)."""
        batch = self.next()
        for img, lab in zip(batch.data[0].asnumpy(), batch.label[0].asnumpy()):
            im = img.transpose(1, 2, 0).copy()
            if std is not None:
                im = im * std
            if mean is not None:
                im = im + mean
            im = np.clip(im, 0, 255).astype(np.uint8) if clip else im.astype(np.uint8)
            H, W = im.shape[:2]
            col = np.asarray(color if color is not None else (255, 0, 0), dtype=np.uint8)
            for row in lab[lab[:, 0] >= 0]:
                x1, y1, x2, y2 = int(row[1] * W), int(row[2] * H), int(row[3] * W), int(row[4] * H)
                x1, x2, y1, y2 = max(0, x1), min(W - 1, x2), max(0, y1), min(H - 1, y2)
                t = thickness
                im[y1:y1 + t, x1:x2] = col; im[max(y2 - t, 0):y2, x1:x2] = col; im[y1:y2, x1:x1 + t] = col; im[y1:y2, max(x2 - t, 0):x2] = col
            yield im
