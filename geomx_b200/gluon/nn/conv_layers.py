"""Convolution / pooling layers.  Parity: ``python/mxnet/gluon/nn/conv_layers.py`` (Conv1D/2D/3D :170-420, Conv1D/2D/3DTranspose :420-700,
Max/AvgPool1D/2D/3D :700-1000, GlobalMax/AvgPool1D/2D/3D, ReflectionPad2D :1170).  Conv2D / MaxPool2D run the native sm_100a kernels on CUDA
(``ops/functional.py``); the 1-D / 3-D / transposed variants are library convolutions."""
from __future__ import annotations

import torch.nn.functional as TF

from ...ndarray import NDArray
from ...ops import functional as OF
from ..block import HybridBlock
from .basic_layers import _bias_init

__all__ = ["Conv1D", "Conv2D", "Conv3D", "Conv1DTranspose", "Conv2DTranspose", "Conv3DTranspose", "MaxPool1D", "MaxPool2D", "MaxPool3D",
           "AvgPool1D", "AvgPool2D", "AvgPool3D", "GlobalAvgPool1D", "GlobalAvgPool2D", "GlobalAvgPool3D", "GlobalMaxPool1D",
           "GlobalMaxPool2D", "GlobalMaxPool3D", "ReflectionPad2D"]


def _pair(v, n=2):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class Conv2D(HybridBlock):
    def __init__(self, channels, kernel_size, strides=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1,
                 layout="NCHW", activation=None, use_bias=True, weight_initializer=None, bias_initializer="zeros",
                 in_channels=0, **kwargs):
        super().__init__(**kwargs)
        assert layout == "NCHW", "only NCHW is supported"
        self._channels, self._in_channels = channels, in_channels
        self._kernel, self._strides, self._padding, self._dilation = _pair(kernel_size), _pair(strides), _pair(padding), _pair(dilation)
        self._groups, self._act = groups, activation
        with self.name_scope():
            self.weight = self.params.get("weight", shape=(channels, in_channels // groups if in_channels else 0) + self._kernel,
                                          init=weight_initializer, allow_deferred_init=True)
            self.bias = self.params.get("bias", shape=(channels,), init=_bias_init(bias_initializer),
                                        allow_deferred_init=True) if use_bias else None

    def _alias(self):
        return "conv"

    def _infer(self, x, *a):
        self.weight.shape = (self._channels, x.shape[1] // self._groups) + self._kernel

    def hybrid_forward(self, F, x, weight, bias=None):
        return NDArray(OF.conv2d(x._t, weight._t, None if bias is None else bias._t, self._strides, self._padding,
                                 self._dilation, self._groups, self._act))

    def __repr__(self):
        return "Conv2D(%s -> %d, kernel_size=%s, stride=%s, %s)" % (
            self.weight.shape[1] if self.weight.shape else None, self._channels, self._kernel, self._strides, self._act or "linear")


class Conv1D(HybridBlock):
    def __init__(self, channels, kernel_size, strides=1, padding=0, dilation=1, groups=1, layout="NCW", activation=None,
                 use_bias=True, weight_initializer=None, bias_initializer="zeros", in_channels=0, **kwargs):
        super().__init__(**kwargs)
        self._channels, self._k, self._s, self._p, self._d, self._groups, self._act = channels, _pair(kernel_size, 1), strides, padding, dilation, groups, activation
        with self.name_scope():
            self.weight = self.params.get("weight", shape=(channels, in_channels // groups if in_channels else 0) + self._k,
                                          init=weight_initializer, allow_deferred_init=True)
            self.bias = self.params.get("bias", shape=(channels,), init=_bias_init(bias_initializer), allow_deferred_init=True) if use_bias else None

    def _alias(self):
        return "conv"

    def _infer(self, x, *a):
        self.weight.shape = (self._channels, x.shape[1] // self._groups) + self._k

    def hybrid_forward(self, F, x, weight, bias=None):
        y = TF.conv1d(x._t, weight._t, None if bias is None else bias._t, self._s, self._p, self._d, self._groups)
        return NDArray(OF._act(y, self._act))


class Conv2DTranspose(HybridBlock):
    def __init__(self, channels, kernel_size, strides=(1, 1), padding=(0, 0), output_padding=(0, 0), dilation=(1, 1),
                 groups=1, layout="NCHW", activation=None, use_bias=True, weight_initializer=None,
                 bias_initializer="zeros", in_channels=0, **kwargs):
        super().__init__(**kwargs)
        self._channels, self._k, self._s, self._p, self._op, self._d, self._groups, self._act = channels, _pair(kernel_size), _pair(strides), _pair(padding), _pair(output_padding), _pair(dilation), groups, activation
        with self.name_scope():
            self.weight = self.params.get("weight", shape=(in_channels, channels // groups) + self._k, init=weight_initializer, allow_deferred_init=True)
            self.bias = self.params.get("bias", shape=(channels,), init=_bias_init(bias_initializer), allow_deferred_init=True) if use_bias else None

    def _alias(self):
        return "conv"

    def _infer(self, x, *a):
        self.weight.shape = (x.shape[1], self._channels // self._groups) + self._k

    def hybrid_forward(self, F, x, weight, bias=None):
        y = TF.conv_transpose2d(x._t, weight._t, None if bias is None else bias._t, self._s, self._p, self._op, self._groups, self._d)
        return NDArray(OF._act(y, self._act))


class _Pool(HybridBlock):
    def __init__(self, pool_size, strides, padding, ceil_mode=False, **kwargs):
        super().__init__(**kwargs)
        self._k = _pair(pool_size); self._s = _pair(strides) if strides is not None else self._k
        self._p = _pair(padding); self._ceil = ceil_mode

    def _alias(self):
        return "pool"


class MaxPool2D(_Pool):
    def __init__(self, pool_size=(2, 2), strides=None, padding=0, layout="NCHW", ceil_mode=False, **kwargs):
        super().__init__(pool_size, strides, padding, ceil_mode, **kwargs)

    def hybrid_forward(self, F, x):
        return NDArray(OF.max_pool2d(x._t, self._k, self._s, self._p, self._ceil))

    def __repr__(self):
        return "MaxPool2D(size=%s, stride=%s, padding=%s)" % (self._k, self._s, self._p)


class AvgPool2D(_Pool):
    def __init__(self, pool_size=(2, 2), strides=None, padding=0, ceil_mode=False, layout="NCHW", count_include_pad=True, **kwargs):
        super().__init__(pool_size, strides, padding, ceil_mode, **kwargs); self._cip = count_include_pad

    def hybrid_forward(self, F, x):
        return NDArray(OF.avg_pool2d(x._t, self._k, self._s, self._p, self._ceil, self._cip))


class MaxPool1D(HybridBlock):
    def __init__(self, pool_size=2, strides=None, padding=0, **kwargs):
        super().__init__(**kwargs); self._k, self._s, self._p = pool_size, strides or pool_size, padding

    def _alias(self):
        return "pool"

    def hybrid_forward(self, F, x):
        return NDArray(TF.max_pool1d(x._t, self._k, self._s, self._p))


class AvgPool1D(MaxPool1D):
    def hybrid_forward(self, F, x):
        return NDArray(TF.avg_pool1d(x._t, self._k, self._s, self._p))


class GlobalAvgPool2D(HybridBlock):
    def _alias(self):
        return "pool"

    def hybrid_forward(self, F, x):
        return NDArray(x._t.mean(dim=(2, 3), keepdim=True))


class GlobalMaxPool2D(HybridBlock):
    def _alias(self):
        return "pool"

    def hybrid_forward(self, F, x):
        return NDArray(x._t.amax(dim=(2, 3), keepdim=True))


class _ConvND(HybridBlock):
    """N-d (transposed) convolution on the library path; weights are laid out like the reference (``(out, in/groups, *k)``, transposed:
    ``(in, out/groups, *k)``)."""
    _nd, _transpose = 3, False

    def __init__(self, channels, kernel_size, strides=1, padding=0, output_padding=0, dilation=1, groups=1, layout=None, activation=None,
                 use_bias=True, weight_initializer=None, bias_initializer="zeros", in_channels=0, **kwargs):
        super().__init__(**kwargs)
        n = self._nd
        self._channels, self._groups, self._act = channels, groups, activation
        self._k, self._s, self._p, self._op, self._d = _pair(kernel_size, n), _pair(strides, n), _pair(padding, n), _pair(output_padding, n), _pair(dilation, n)
        wshape = ((in_channels, channels // groups) if self._transpose else (channels, in_channels // groups if in_channels else 0)) + self._k
        with self.name_scope():
            self.weight = self.params.get("weight", shape=wshape, init=weight_initializer, allow_deferred_init=True)
            self.bias = self.params.get("bias", shape=(channels,), init=_bias_init(bias_initializer), allow_deferred_init=True) if use_bias else None

    def _alias(self):
        return "conv"

    def _infer(self, x, *a):
        self.weight.shape = ((x.shape[1], self._channels // self._groups) if self._transpose else (self._channels, x.shape[1] // self._groups)) + self._k

    def hybrid_forward(self, F, x, weight, bias=None):
        b = None if bias is None else bias._t
        if self._transpose:
            fn = (TF.conv_transpose1d, TF.conv_transpose2d, TF.conv_transpose3d)[self._nd - 1]
            y = fn(x._t, weight._t, b, self._s, self._p, self._op, self._groups, self._d)
        else:
            fn = (TF.conv1d, TF.conv2d, TF.conv3d)[self._nd - 1]
            y = fn(x._t, weight._t, b, self._s, self._p, self._d, self._groups)
        return NDArray(OF._act(y, self._act))


class Conv3D(_ConvND):
    _nd, _transpose = 3, False


class Conv1DTranspose(_ConvND):
    _nd, _transpose = 1, True


class Conv3DTranspose(_ConvND):
    _nd, _transpose = 3, True


class MaxPool3D(_Pool):
    def __init__(self, pool_size=(2, 2, 2), strides=None, padding=0, ceil_mode=False, layout="NCDHW", **kwargs):
        HybridBlock.__init__(self, **kwargs)
        self._k = _pair(pool_size, 3); self._s = _pair(strides, 3) if strides is not None else self._k; self._p = _pair(padding, 3); self._ceil = ceil_mode

    def hybrid_forward(self, F, x):
        return NDArray(TF.max_pool3d(x._t, self._k, self._s, self._p, ceil_mode=self._ceil))


class AvgPool3D(MaxPool3D):
    def __init__(self, pool_size=(2, 2, 2), strides=None, padding=0, ceil_mode=False, layout="NCDHW", count_include_pad=True, **kwargs):
        super().__init__(pool_size, strides, padding, ceil_mode, **kwargs); self._cip = count_include_pad

    def hybrid_forward(self, F, x):
        return NDArray(TF.avg_pool3d(x._t, self._k, self._s, self._p, ceil_mode=self._ceil, count_include_pad=self._cip))


def _global_pool(name, reduce, dims):
    def hybrid_forward(self, F, x):
        return NDArray(getattr(x._t, reduce)(dim=dims, keepdim=True))
    return type(name, (HybridBlock,), {"hybrid_forward": hybrid_forward, "_alias": lambda self: "pool", "__init__": lambda self, layout=None, **kw: HybridBlock.__init__(self, **kw)})


GlobalAvgPool1D = _global_pool("GlobalAvgPool1D", "mean", (2,))
GlobalAvgPool3D = _global_pool("GlobalAvgPool3D", "mean", (2, 3, 4))
GlobalMaxPool1D = _global_pool("GlobalMaxPool1D", "amax", (2,))
GlobalMaxPool3D = _global_pool("GlobalMaxPool3D", "amax", (2, 3, 4))


class ReflectionPad2D(HybridBlock):
    """Reflection padding of the two spatial axes; ``padding`` is an int or ``(left, right, top, bottom)`` / MXNet's 8-tuple pad_width."""

    def __init__(self, padding=0, **kwargs):
        super().__init__(**kwargs)
        if isinstance(padding, int):
            padding = (padding,) * 4
        elif len(padding) == 8:                                # (0,0,0,0, top, bottom, left, right)
            padding = (padding[6], padding[7], padding[4], padding[5])
        self._pad = tuple(int(p) for p in padding)

    def hybrid_forward(self, F, x):
        return NDArray(TF.pad(x._t, self._pad, mode="reflect"))
