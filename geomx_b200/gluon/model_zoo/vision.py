"""Vision models (parity: ``python/mxnet/gluon/model_zoo/vision/{alexnet,vgg,resnet,squeezenet,mobilenet}.py``; same layer recipes, built from
``gluon.nn`` so that Conv/Dense/Pool/BatchNorm run through the native kernels on CUDA)."""
from __future__ import annotations

from ...base import MXNetError
from .. import nn
from ..block import HybridBlock

__all__ = ["get_model", "AlexNet", "alexnet", "VGG", "vgg11", "vgg13", "vgg16", "vgg11_bn", "ResNetV1", "resnet18_v1", "resnet34_v1",
           "SqueezeNet", "squeezenet1_1", "MobileNet", "mobilenet1_0", "mobilenet0_5", "LeNet", "lenet", "MLP", "mlp"]


class LeNet(HybridBlock):
    """The GeoMX demo CNN family (examples/cnn.py: Conv-Pool-Conv-Pool-Dense-Dense-Dense)."""

    def __init__(self, classes=10, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            self.features.add(nn.Conv2D(16, 5, activation="relu"), nn.MaxPool2D(2, 2), nn.Conv2D(32, 5, activation="relu"), nn.MaxPool2D(2, 2), nn.Flatten(),
                              nn.Dense(256, activation="relu"), nn.Dense(128, activation="relu"))
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class MLP(HybridBlock):
    def __init__(self, hidden=(512, 256), classes=10, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            self.features.add(nn.Flatten())
            for h in hidden:
                self.features.add(nn.Dense(h, activation="relu"))
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class AlexNet(HybridBlock):
    def __init__(self, classes=1000, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            f = self.features = nn.HybridSequential(prefix="")
            f.add(nn.Conv2D(64, 11, 4, 2, activation="relu"), nn.MaxPool2D(3, 2), nn.Conv2D(192, 5, padding=2, activation="relu"), nn.MaxPool2D(3, 2),
                  nn.Conv2D(384, 3, padding=1, activation="relu"), nn.Conv2D(256, 3, padding=1, activation="relu"),
                  nn.Conv2D(256, 3, padding=1, activation="relu"), nn.MaxPool2D(3, 2), nn.Flatten(), nn.Dense(4096, activation="relu"), nn.Dropout(0.5),
                  nn.Dense(4096, activation="relu"), nn.Dropout(0.5))
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class VGG(HybridBlock):
    def __init__(self, layers, filters, classes=1000, batch_norm=False, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            f = self.features = nn.HybridSequential(prefix="")
            for n, c in zip(layers, filters):
                for _ in range(n):
                    f.add(nn.Conv2D(c, 3, padding=1))
                    if batch_norm:
                        f.add(nn.BatchNorm())
                    f.add(nn.Activation("relu"))
                f.add(nn.MaxPool2D(2, 2))
            f.add(nn.Flatten(), nn.Dense(4096, activation="relu"), nn.Dropout(0.5), nn.Dense(4096, activation="relu"), nn.Dropout(0.5))
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class BasicBlockV1(HybridBlock):
    def __init__(self, channels, stride, downsample=False, in_channels=0, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            self.body = nn.HybridSequential(prefix="")
            self.body.add(nn.Conv2D(channels, 3, stride, 1, use_bias=False, in_channels=in_channels), nn.BatchNorm(), nn.Activation("relu"),
                          nn.Conv2D(channels, 3, 1, 1, use_bias=False, in_channels=channels), nn.BatchNorm())
            self.downsample = None
            if downsample:
                self.downsample = nn.HybridSequential(prefix="")
                self.downsample.add(nn.Conv2D(channels, 1, stride, use_bias=False, in_channels=in_channels), nn.BatchNorm())

    def hybrid_forward(self, F, x):
        res = x if self.downsample is None else self.downsample(x)
        return (self.body(x) + res).relu()


class ResNetV1(HybridBlock):
    def __init__(self, layers, channels, classes=1000, thumbnail=False, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            f = self.features = nn.HybridSequential(prefix="")
            if thumbnail:
                f.add(nn.Conv2D(channels[0], 3, 1, 1, use_bias=False))
            else:
                f.add(nn.Conv2D(channels[0], 7, 2, 3, use_bias=False), nn.BatchNorm(), nn.Activation("relu"), nn.MaxPool2D(3, 2, 1))
            for i, n in enumerate(layers):
                stride = 1 if i == 0 else 2
                f.add(BasicBlockV1(channels[i + 1], stride, channels[i + 1] != channels[i] or stride != 1, in_channels=channels[i]))
                for _ in range(n - 1):
                    f.add(BasicBlockV1(channels[i + 1], 1, False, in_channels=channels[i + 1]))
            f.add(nn.GlobalAvgPool2D(), nn.Flatten())
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class _Fire(HybridBlock):
    def __init__(self, squeeze, e1, e3, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            self.squeeze = nn.Conv2D(squeeze, 1, activation="relu")
            self.e1 = nn.Conv2D(e1, 1, activation="relu")
            self.e3 = nn.Conv2D(e3, 3, padding=1, activation="relu")

    def hybrid_forward(self, F, x):
        from ... import ndarray as nd
        s = self.squeeze(x)
        return nd.concat(self.e1(s), self.e3(s), dim=1)


class SqueezeNet(HybridBlock):
    def __init__(self, classes=1000, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            f = self.features = nn.HybridSequential(prefix="")
            f.add(nn.Conv2D(64, 3, 2, activation="relu"), nn.MaxPool2D(3, 2), _Fire(16, 64, 64), _Fire(16, 64, 64), nn.MaxPool2D(3, 2),
                  _Fire(32, 128, 128), _Fire(32, 128, 128), nn.MaxPool2D(3, 2), _Fire(48, 192, 192), _Fire(48, 192, 192), _Fire(64, 256, 256),
                  _Fire(64, 256, 256), nn.Dropout(0.5))
            self.output = nn.HybridSequential(prefix="")
            self.output.add(nn.Conv2D(classes, 1, activation="relu"), nn.GlobalAvgPool2D(), nn.Flatten())

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class MobileNet(HybridBlock):
    def __init__(self, multiplier=1.0, classes=1000, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            f = self.features = nn.HybridSequential(prefix="")

            def conv(c, k=1, s=1, p=0, g=1, cin=0):
                f.add(nn.Conv2D(c, k, s, p, groups=g, use_bias=False, in_channels=cin), nn.BatchNorm(), nn.Activation("relu"))
            c0 = int(32 * multiplier)
            conv(c0, 3, 2, 1, cin=0)
            dw = [int(x * multiplier) for x in [32, 64] + [128] * 2 + [256] * 2 + [512] * 6 + [1024]]
            pw = [int(x * multiplier) for x in [64] + [128] * 2 + [256] * 2 + [512] * 6 + [1024] * 2]
            strides = [1, 2] * 3 + [1] * 5 + [2, 1]
            for d, c, s in zip(dw, pw, strides):
                conv(d, 3, s, 1, g=d, cin=d)
                conv(c, cin=d)
            f.add(nn.GlobalAvgPool2D(), nn.Flatten())
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


def lenet(**kw): return LeNet(**kw)
def mlp(**kw): return MLP(**kw)
def alexnet(**kw): return AlexNet(**kw)
def vgg11(**kw): return VGG([1, 1, 2, 2, 2], [64, 128, 256, 512, 512], **kw)
def vgg13(**kw): return VGG([2, 2, 2, 2, 2], [64, 128, 256, 512, 512], **kw)
def vgg16(**kw): return VGG([2, 2, 3, 3, 3], [64, 128, 256, 512, 512], **kw)
def vgg11_bn(**kw): return VGG([1, 1, 2, 2, 2], [64, 128, 256, 512, 512], batch_norm=True, **kw)
def resnet18_v1(**kw): return ResNetV1([2, 2, 2, 2], [64, 64, 128, 256, 512], **kw)
def resnet34_v1(**kw): return ResNetV1([3, 4, 6, 3], [64, 64, 128, 256, 512], **kw)
def squeezenet1_1(**kw): return SqueezeNet(**kw)
def mobilenet1_0(**kw): return MobileNet(1.0, **kw)
def mobilenet0_5(**kw): return MobileNet(0.5, **kw)


_models = {"lenet": lenet, "mlp": mlp, "alexnet": alexnet, "vgg11": vgg11, "vgg13": vgg13, "vgg16": vgg16, "vgg11_bn": vgg11_bn,
           "resnet18_v1": resnet18_v1, "resnet34_v1": resnet34_v1, "squeezenet1.1": squeezenet1_1, "mobilenet1.0": mobilenet1_0,
           "mobilenet0.5": mobilenet0_5}


def get_model(name, pretrained=False, **kwargs):
    if pretrained:
        raise MXNetError("pretrained weights are not available offline; load a .params file with net.load_parameters")
    name = name.lower()
    if name not in _models:
        raise MXNetError("Model %s is not supported. Available: %s" % (name, ", ".join(sorted(_models))))
    return _models[name](**kwargs)
