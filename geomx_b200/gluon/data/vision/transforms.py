"""Vision transforms.  Parity: ``python/mxnet/gluon/data/vision/transforms.py`` (Compose, Cast, ToTensor: HWC uint8
→ CHW float32/255, Normalize, Resize, CenterCrop, RandomFlipLeftRight)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ....ndarray import NDArray
from ...block import Block
from ...nn import Sequential

__all__ = ["Compose", "Cast", "ToTensor", "Normalize", "Resize", "CenterCrop", "RandomFlipLeftRight", "RandomSaturation", "RandomHue", "RandomColorJitter", "RandomLighting"]


class Compose(Sequential):
    def __init__(self, transforms):
        super().__init__()
        for t in transforms:
            self.add(t)


class Cast(Block):
    def __init__(self, dtype="float32"):
        super().__init__(); self._dtype = dtype

    def forward(self, x):
        return x.astype(self._dtype)


class ToTensor(Block):
    def forward(self, x):
        t = x._t
        return NDArray(t.permute(2, 0, 1).to(torch.float32).div_(255.0)) if t.dim() == 3 else \
            NDArray(t.permute(0, 3, 1, 2).to(torch.float32).div_(255.0))


class Normalize(Block):
    def __init__(self, mean, std):
        super().__init__(); self._mean, self._std = mean, std

    def forward(self, x):
        t = x._t
        m = torch.as_tensor(self._mean, dtype=t.dtype).reshape(-1, 1, 1)
        s = torch.as_tensor(self._std, dtype=t.dtype).reshape(-1, 1, 1)
        return NDArray((t - m) / s)


class Resize(Block):
    def __init__(self, size, keep_ratio=False, interpolation=1):
        super().__init__()
        self._size = (size, size) if isinstance(size, int) else tuple(size)

    def forward(self, x):
        t = x._t  # HWC
        h, w = self._size[1], self._size[0]
        if t.shape[0] == h and t.shape[1] == w:
            return x
        y = F.interpolate(t.permute(2, 0, 1).unsqueeze(0).float(), size=(h, w), mode="bilinear", align_corners=False)
        return NDArray(y.squeeze(0).permute(1, 2, 0).round().clamp(0, 255).to(t.dtype))


class CenterCrop(Block):
    def __init__(self, size, interpolation=1):
        super().__init__(); self._size = (size, size) if isinstance(size, int) else tuple(size)

    def forward(self, x):
        t = x._t; w, h = self._size
        y0 = max(0, (t.shape[0] - h) // 2); x0 = max(0, (t.shape[1] - w) // 2)
        return NDArray(t[y0:y0 + h, x0:x0 + w])


class RandomFlipLeftRight(Block):
    def forward(self, x):
        return NDArray(x._t.flip(1)) if torch.rand(()) < 0.5 else x


class RandomFlipTopBottom(Block):
    def forward(self, x):
        import random
        return NDArray(x._t.flip(0)) if random.random() < 0.5 else x


class RandomResizedCrop(Block):
    """Random area/aspect crop resized to ``size`` (transforms.py RandomResizedCrop), input HWC."""

    def __init__(self, size, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), interpolation=1):
        super().__init__()
        self._size = (size, size) if isinstance(size, int) else tuple(size)
        self._scale, self._ratio = scale, ratio

    def forward(self, x):
        import math
        import random
        import torch.nn.functional as TF
        h, w = x.shape[0], x.shape[1]
        for _ in range(10):
            area = random.uniform(*self._scale) * h * w
            ar = math.exp(random.uniform(math.log(self._ratio[0]), math.log(self._ratio[1])))
            cw, ch = int(round(math.sqrt(area * ar))), int(round(math.sqrt(area / ar)))
            if 0 < cw <= w and 0 < ch <= h:
                x0, y0 = random.randint(0, w - cw), random.randint(0, h - ch)
                break
        else:
            cw, ch, x0, y0 = w, h, 0, 0
        crop = x._t[y0:y0 + ch, x0:x0 + cw].permute(2, 0, 1).unsqueeze(0).float()
        out = TF.interpolate(crop, size=(self._size[1], self._size[0]), mode="bilinear", align_corners=False)[0].permute(1, 2, 0)
        return NDArray(out.to(x._t.dtype) if x._t.dtype.is_floating_point else out.round().clamp(0, 255).to(x._t.dtype))


class RandomBrightness(Block):
    def __init__(self, brightness):
        super().__init__(); self._b = brightness

    def forward(self, x):
        import random
        return NDArray(x._t.float() * (1.0 + random.uniform(-self._b, self._b)))


class RandomContrast(Block):
    def __init__(self, contrast):
        super().__init__(); self._c = contrast

    def forward(self, x):
        import random
        t = x._t.float()
        alpha = 1.0 + random.uniform(-self._c, self._c)
        return NDArray(t * alpha + t.mean() * (1 - alpha))


class RandomSaturation(Block):
    def __init__(self, saturation):
        super().__init__(); self._s = saturation

    def forward(self, x):
        from .... import image
        return image.SaturationJitterAug(self._s)(x)


class RandomHue(Block):
    def __init__(self, hue):
        super().__init__(); self._h = hue

    def forward(self, x):
        from .... import image
        return image.HueJitterAug(self._h)(x)


class RandomColorJitter(Block):
    """Brightness / contrast / saturation / hue jitter in random order (transforms.py RandomColorJitter)."""

    def __init__(self, brightness=0, contrast=0, saturation=0, hue=0):
        super().__init__()
        from .... import image
        self._augs = [a for a in ([image.BrightnessJitterAug(brightness)] if brightness else []) + ([image.ContrastJitterAug(contrast)] if contrast else [])
                      + ([image.SaturationJitterAug(saturation)] if saturation else []) + ([image.HueJitterAug(hue)] if hue else [])]

    def forward(self, x):
        import random
        order = list(self._augs); random.shuffle(order)
        for a in order:
            x = a(x)
        return x


class RandomLighting(Block):
    """AlexNet-style PCA noise with standard deviation ``alpha``."""

    def __init__(self, alpha):
        super().__init__(); self._alpha = alpha

    def forward(self, x):
        import numpy as np
        from .... import image
        return image.LightingAug(self._alpha, np.array([55.46, 4.794, 1.148]),
                                 np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]]))(x)
