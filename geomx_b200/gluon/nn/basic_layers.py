"""Basic Gluon layers.  Parity: ``python/mxnet/gluon/nn/basic_layers.py`` (Sequential :32, HybridSequential,
Dense :142, Dropout, BatchNorm, Embedding, Flatten, InstanceNorm-free subset, LayerNorm, Lambda, HybridLambda)
and ``activations.py`` (Activation, LeakyReLU, PReLU-free, ELU, SELU, Swish)."""
from __future__ import annotations

import torch

from ... import autograd
from ...ndarray import NDArray
from ...ops import functional as OF
from ..block import Block, HybridBlock

__all__ = ["Sequential", "HybridSequential", "Dense", "Dropout", "BatchNorm", "Embedding", "Flatten", "LayerNorm",
           "Lambda", "HybridLambda", "Activation", "LeakyReLU", "ELU", "SELU", "Swish", "PReLU", "GELU", "InstanceNorm"]


class Sequential(Block):
    def __init__(self, prefix=None, params=None):
        super().__init__(prefix, params)

    def add(self, *blocks):
        for b in blocks:
            self.register_child(b)

    def forward(self, x):
        for b in self._children.values():
            x = b(x)
        return x

    def __getitem__(self, key):
        layers = list(self._children.values())[key]
        if isinstance(layers, list):
            net = type(self)(prefix=self._prefix)
            with net.name_scope():
                net.add(*layers)
            return net
        return layers

    def __len__(self):
        return len(self._children)

    def __iter__(self):
        return iter(self._children.values())


class HybridSequential(Sequential, HybridBlock):
    def __init__(self, prefix=None, params=None):
        HybridBlock.__init__(self, prefix, params)

    def forward(self, x):
        return Sequential.forward(self, x)


class Dense(HybridBlock):
    """``y = act(x·Wᵀ + b)``; weight shape (units, in_units); ``in_units=0`` → deferred."""

    def __init__(self, units, activation=None, use_bias=True, flatten=True, dtype="float32",
                 weight_initializer=None, bias_initializer="zeros", in_units=0, **kwargs):
        super().__init__(**kwargs)
        self._flatten, self._units, self._in_units, self._act = flatten, units, in_units, activation
        with self.name_scope():
            self.weight = self.params.get("weight", shape=(units, in_units), init=weight_initializer, dtype=dtype,
                                          allow_deferred_init=True)
            self.bias = self.params.get("bias", shape=(units,), init=_bias_init(bias_initializer), dtype=dtype,
                                        allow_deferred_init=True) if use_bias else None

    def _infer(self, x, *a):
        in_units = 1
        if self._flatten:
            for s in x.shape[1:]:
                in_units *= s
        else:
            in_units = x.shape[-1]
        self.weight.shape = (self._units, in_units)

    def hybrid_forward(self, F, x, weight, bias=None):
        return NDArray(OF.dense(x._t, weight._t, None if bias is None else bias._t, self._act, self._flatten))

    def __repr__(self):
        return "Dense(%s -> %d, %s)" % (self.weight.shape[1] if self.weight.shape else None, self._units, self._act or "linear")


def _bias_init(b):
    from ... import initializer
    if b is None or isinstance(b, initializer.Initializer):
        return b
    return initializer.create(b if b != "zeros" else "zero") if isinstance(b, str) else b


class Activation(HybridBlock):
    def __init__(self, activation, **kwargs):
        self._act_type = activation
        super().__init__(**kwargs)

    def _alias(self):
        return self._act_type

    def hybrid_forward(self, F, x):
        return NDArray(OF.activation(x._t, self._act_type))


class LeakyReLU(HybridBlock):
    def __init__(self, alpha, **kwargs):
        super().__init__(**kwargs); self._alpha = alpha

    def hybrid_forward(self, F, x):
        return NDArray(torch.nn.functional.leaky_relu(x._t, self._alpha))


class ELU(HybridBlock):
    def __init__(self, alpha=1.0, **kwargs):
        super().__init__(**kwargs); self._alpha = alpha

    def hybrid_forward(self, F, x):
        return NDArray(torch.nn.functional.elu(x._t, self._alpha))


class SELU(HybridBlock):
    def hybrid_forward(self, F, x):
        return NDArray(torch.selu(x._t))


class Swish(HybridBlock):
    def __init__(self, beta=1.0, **kwargs):
        super().__init__(**kwargs); self._beta = beta

    def hybrid_forward(self, F, x):
        return NDArray(x._t * torch.sigmoid(self._beta * x._t))


class Dropout(HybridBlock):
    def __init__(self, rate, axes=(), **kwargs):
        super().__init__(**kwargs); self._rate = rate

    def hybrid_forward(self, F, x):
        return NDArray(OF.dropout(x._t, self._rate, autograd.is_training()))


class BatchNorm(HybridBlock):
    def __init__(self, axis=1, momentum=0.9, epsilon=1e-5, center=True, scale=True, use_global_stats=False,
                 beta_initializer="zeros", gamma_initializer="ones", running_mean_initializer="zeros",
                 running_variance_initializer="ones", in_channels=0, **kwargs):
        super().__init__(**kwargs)
        self._axis, self._momentum, self._eps, self._use_global = axis, momentum, epsilon, use_global_stats
        shape = (in_channels,)
        with self.name_scope():
            self.gamma = self.params.get("gamma", grad_req="write" if scale else "null", shape=shape,
                                         init=_bias_init("one"), allow_deferred_init=True, differentiable=scale)
            self.beta = self.params.get("beta", grad_req="write" if center else "null", shape=shape,
                                        init=_bias_init("zero"), allow_deferred_init=True, differentiable=center)
            self.running_mean = self.params.get("running_mean", grad_req="null", shape=shape, init=_bias_init("zero"),
                                                allow_deferred_init=True, differentiable=False)
            self.running_var = self.params.get("running_var", grad_req="null", shape=shape, init=_bias_init("one"),
                                               allow_deferred_init=True, differentiable=False)

    def _infer(self, x, *a):
        c = x.shape[self._axis]
        for p in (self.gamma, self.beta, self.running_mean, self.running_var):
            p.shape = (c,)

    def hybrid_forward(self, F, x, gamma, beta, running_mean, running_var):
        training = autograd.is_training() and not self._use_global
        return NDArray(OF.batch_norm(x._t, gamma._t, beta._t, running_mean._t, running_var._t, training,
                                     self._momentum, self._eps, self._axis))


class LayerNorm(HybridBlock):
    def __init__(self, axis=-1, epsilon=1e-5, center=True, scale=True, in_channels=0, **kwargs):
        super().__init__(**kwargs)
        self._axis, self._eps = axis, epsilon
        with self.name_scope():
            self.gamma = self.params.get("gamma", shape=(in_channels,), init=_bias_init("one"), allow_deferred_init=True)
            self.beta = self.params.get("beta", shape=(in_channels,), init=_bias_init("zero"), allow_deferred_init=True)

    def _infer(self, x, *a):
        self.gamma.shape = self.beta.shape = (x.shape[self._axis],)

    def hybrid_forward(self, F, x, gamma, beta):
        return NDArray(OF.layer_norm(x._t, gamma._t, beta._t, self._axis, self._eps))


class Embedding(HybridBlock):
    def __init__(self, input_dim, output_dim, dtype="float32", weight_initializer=None, sparse_grad=False, **kwargs):
        super().__init__(**kwargs)
        with self.name_scope():
            self.weight = self.params.get("weight", shape=(input_dim, output_dim), init=weight_initializer, dtype=dtype)

    def hybrid_forward(self, F, x, weight):
        return NDArray(torch.nn.functional.embedding(x._t.long(), weight._t))


class Flatten(HybridBlock):
    def hybrid_forward(self, F, x):
        return NDArray(OF.flatten(x._t))


class Lambda(Block):
    def __init__(self, function, prefix=None):
        super().__init__(prefix=prefix)
        from ... import ndarray as nd
        self._func_impl = getattr(nd, function) if isinstance(function, str) else function

    def forward(self, *args):
        return self._func_impl(*args)


class HybridLambda(HybridBlock):
    def __init__(self, function, prefix=None):
        super().__init__(prefix=prefix)
        from ... import ndarray as nd
        self._func = (lambda F, *a: getattr(nd, function)(*a)) if isinstance(function, str) else function

    def hybrid_forward(self, F, x, *args):
        return self._func(F, x, *args)


class PReLU(HybridBlock):
    """``max(0, x) + alpha * min(0, x)`` with a learned ``alpha`` (one value, or one per channel with ``in_channels``)
    (``gluon/nn/activations.py:87-125``)."""

    def __init__(self, alpha_initializer=None, in_channels=1, **kwargs):
        super().__init__(**kwargs)
        from ... import initializer
        with self.name_scope():
            self.alpha = self.params.get("alpha", shape=(in_channels,), init=alpha_initializer or initializer.Constant(0.25))

    def hybrid_forward(self, F, x, alpha):
        a = alpha._t
        if a.numel() > 1:
            a = a.view([1, -1] + [1] * (x._t.dim() - 2))
        return NDArray(torch.where(x._t >= 0, x._t, a * x._t))


class GELU(HybridBlock):
    def hybrid_forward(self, F, x):
        return NDArray(torch.nn.functional.gelu(x._t))


class InstanceNorm(HybridBlock):
    """Per-sample, per-channel normalisation over the spatial axes (``gluon/nn/basic_layers.py:455-540``)."""

    def __init__(self, axis=1, epsilon=1e-5, center=True, scale=False, beta_initializer="zeros", gamma_initializer="ones", in_channels=0, **kwargs):
        super().__init__(**kwargs)
        self._axis, self._eps = axis, epsilon
        with self.name_scope():
            self.gamma = self.params.get("gamma", grad_req="write" if scale else "null", shape=(in_channels,), init=_bias_init("one"), allow_deferred_init=True)
            self.beta = self.params.get("beta", grad_req="write" if center else "null", shape=(in_channels,), init=_bias_init("zero"), allow_deferred_init=True)

    def _infer(self, x, *a):
        self.gamma.shape = self.beta.shape = (x.shape[self._axis],)

    def hybrid_forward(self, F, x, gamma, beta):
        t = x._t if self._axis == 1 else x._t.movedim(self._axis, 1)
        y = torch.nn.functional.instance_norm(t, weight=gamma._t, bias=beta._t, eps=self._eps)
        return NDArray(y if self._axis == 1 else y.movedim(1, self._axis))
