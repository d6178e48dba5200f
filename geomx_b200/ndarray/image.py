"""``mx.nd.image`` — image operators on HWC (or NHWC) arrays (reference: ``python/mxnet/ndarray/image.py`` over the ``_image_*`` operators of
``src/operator/image/image_random-inl.h``: to_tensor, normalize, flips, brightness / contrast / saturation / hue jitter, PCA lighting)."""
import numpy as np
import torch

from .ndarray import NDArray

__all__ = ["to_tensor", "normalize", "flip_left_right", "flip_top_bottom", "random_flip_left_right", "random_flip_top_bottom", "random_brightness",
           "random_contrast", "random_saturation", "random_hue", "random_color_jitter", "adjust_lighting", "random_lighting", "resize", "crop"]

_GRAY = (0.299, 0.587, 0.114)


def _w(t):
    return NDArray(t)


def to_tensor(data):
    """HWC (or NHWC) uint8 in [0, 255] -> CHW (NCHW) float32 in [0, 1)."""
    t = data._t
    t = t.permute(2, 0, 1) if t.dim() == 3 else t.permute(0, 3, 1, 2)
    return _w(t.to(torch.float32) / 255.0)


def normalize(data, mean=0.0, std=1.0):
    """Per-channel ``(x - mean) / std`` of a CHW (NCHW) float image."""
    t = data._t
    shape = (-1, 1, 1)
    m = torch.as_tensor(mean, dtype=t.dtype, device=t.device).reshape(shape) if np.ndim(mean) else mean
    s = torch.as_tensor(std, dtype=t.dtype, device=t.device).reshape(shape) if np.ndim(std) else std
    return _w((t - m) / s)


def flip_left_right(data):
    return _w(data._t.flip(-2))          # HWC: width is the second-to-last axis


def flip_top_bottom(data):
    return _w(data._t.flip(-3))


def random_flip_left_right(data):
    return flip_left_right(data) if np.random.rand() < 0.5 else data


def random_flip_top_bottom(data):
    return flip_top_bottom(data) if np.random.rand() < 0.5 else data


def _blend(t, other, alpha):
    out = t.to(torch.float32) * alpha + other * (1.0 - alpha)
    return out.clamp(0, 255).to(t.dtype) if not t.dtype.is_floating_point else out.to(t.dtype)


def _gray(t):
    coef = torch.tensor(_GRAY, dtype=torch.float32, device=t.device)
    return (t.to(torch.float32) * coef).sum(-1, keepdim=True)


def random_brightness(data, min_factor, max_factor):
    return _w(_blend(data._t, 0.0, float(np.random.uniform(min_factor, max_factor))))


def random_contrast(data, min_factor, max_factor):
    t = data._t
    return _w(_blend(t, _gray(t).mean(), float(np.random.uniform(min_factor, max_factor))))


def random_saturation(data, min_factor, max_factor):
    t = data._t
    return _w(_blend(t, _gray(t), float(np.random.uniform(min_factor, max_factor))))


def random_hue(data, min_factor, max_factor):
    """Rotate the hue by a random angle: ``alpha`` in ``[min_factor, max_factor]`` maps to a rotation of ``alpha * pi`` in YIQ space."""
    t = data._t
    alpha = float(np.random.uniform(min_factor, max_factor))
    u, w = np.cos(alpha * np.pi), np.sin(alpha * np.pi)
    bt = np.array([[1.0, 0.0, 0.0], [0.0, u, -w], [0.0, w, u]])
    tyiq = np.array([[0.299, 0.587, 0.114], [0.596, -0.274, -0.321], [0.211, -0.523, 0.311]])
    ityiq = np.array([[1.0, 0.956, 0.621], [1.0, -0.272, -0.647], [1.0, -1.107, 1.705]])
    m = torch.tensor(np.dot(np.dot(ityiq, bt), tyiq).T, dtype=torch.float32, device=t.device)
    out = t.to(torch.float32) @ m
    return _w(out.clamp(0, 255).to(t.dtype) if not t.dtype.is_floating_point else out.to(t.dtype))


def random_color_jitter(data, brightness=0.0, contrast=0.0, saturation=0.0, hue=0.0):
    ops = []
    if brightness > 0: ops.append(lambda x: random_brightness(x, 1 - brightness, 1 + brightness))
    if contrast > 0: ops.append(lambda x: random_contrast(x, 1 - contrast, 1 + contrast))
    if saturation > 0: ops.append(lambda x: random_saturation(x, 1 - saturation, 1 + saturation))
    if hue > 0: ops.append(lambda x: random_hue(x, -hue, hue))
    for i in np.random.permutation(len(ops)):
        data = ops[i](data)
    return data


_EIGVAL = np.array([55.46, 4.794, 1.148])
_EIGVEC = np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]])


def adjust_lighting(data, alpha):
    """AlexNet-style PCA lighting: add ``eigvec @ (alpha * eigval)`` to every pixel."""
    t = data._t
    rgb = torch.tensor(np.dot(_EIGVEC * np.asarray(alpha, dtype=np.float64), _EIGVAL), dtype=torch.float32, device=t.device)
    out = t.to(torch.float32) + rgb
    return _w(out.clamp(0, 255).to(t.dtype) if not t.dtype.is_floating_point else out.to(t.dtype))


def random_lighting(data, alpha_std=0.05):
    return adjust_lighting(data, np.random.normal(0, alpha_std, size=(3,)))


def resize(data, size, keep_ratio=False, interp=1):
    """Resize an HWC / NHWC image to ``size`` = (width, height) (an int: both, or the shorter side with ``keep_ratio``)."""
    t = data._t
    h, w = t.shape[-3], t.shape[-2]
    if isinstance(size, int):
        if keep_ratio:
            new_w, new_h = (size, int(h * size / w)) if w < h else (int(w * size / h), size)
        else:
            new_w = new_h = size
    else:
        new_w, new_h = size
    x = (t.permute(2, 0, 1)[None] if t.dim() == 3 else t.permute(0, 3, 1, 2)).to(torch.float32)
    mode = {0: "nearest", 1: "bilinear", 2: "bicubic", 3: "area"}.get(interp, "bilinear")
    y = torch.nn.functional.interpolate(x, size=(new_h, new_w), mode=mode, **({} if mode in ("nearest", "area") else {"align_corners": False}))
    y = y[0].permute(1, 2, 0) if t.dim() == 3 else y.permute(0, 2, 3, 1)
    return _w(y.round().clamp(0, 255).to(t.dtype) if not t.dtype.is_floating_point else y.to(t.dtype))


def crop(data, x, y, width, height):
    return _w(data._t[..., y:y + height, x:x + width, :])
