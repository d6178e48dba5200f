"""Weight initialisers (``mx.init``).

Parity: ``python/mxnet/initializer.py`` — name-pattern dispatch of ``Initializer.__call__``
(weight/bias/gamma/beta/moving_*), ``Xavier`` (rnd_type, factor_type, magnitude), ``Uniform``,
``Normal``, ``Constant``, ``Zero``, ``One``, ``Orthogonal``, ``MSRAPrelu``, ``Bilinear``, ``Mixed``,
``Load``, and the ``@register`` / ``create`` registry."""
from __future__ import annotations

import math
import re

import numpy as np
import torch

from .base import MXNetError

__all__ = ["Initializer", "Xavier", "Uniform", "Normal", "Constant", "Zero", "One", "Orthogonal", "LSTMBias", "FusedRNN",
           "MSRAPrelu", "Bilinear", "Mixed", "Load", "InitDesc", "register", "create"]

_registry = {}


def register(klass):
    _registry[klass.__name__.lower()] = klass
    return klass


def create(name, **kwargs):
    if isinstance(name, Initializer):
        return name
    if name is None:
        return Uniform()
    k = str(name).lower()
    if k not in _registry:
        raise MXNetError("unknown initializer %s" % name)
    return _registry[k](**kwargs)


class InitDesc(str):
    def __new__(cls, name, attrs=None, global_init=None):
        ret = super().__new__(cls, name)
        ret.attrs = attrs or {}
        ret.global_init = global_init
        return ret


class Initializer:
    def __init__(self, **kwargs):
        self._kwargs = kwargs

    def dumps(self):
        import json
        return json.dumps([self.__class__.__name__.lower(), self._kwargs])

    def set_verbosity(self, verbose=False, print_func=None):
        """Log a statistic of every array right after it was initialised (``print_func(array) -> str``; default: its L2 norm / sqrt(size))."""
        self._verbose = bool(verbose)
        self._print_func = print_func or (lambda x: str(float(x.norm()) / max(1.0, float(x.numel()) ** 0.5)))
        return self

    def _verbose_print(self, desc, init, arr):
        if getattr(self, "_verbose", False):
            import logging
            logging.info("Initialized %s as %s: %s", desc, init, self._print_func(arr))

    def __call__(self, desc, arr):
        name = str(desc)
        t = arr._t if hasattr(arr, "_t") else arr
        t = t.detach()
        own = getattr(desc, "attrs", None) and desc.attrs.get("__init__")
        if own:                                   # the variable names its own initializer (mx.sym.Variable(..., init=...)): it wins
            import json
            klass, kwargs = json.loads(own)
            if getattr(desc, "global_init", None) is None and isinstance(desc, InitDesc):
                desc.global_init = self          # composite initializers (FusedRNN) fill their blocks with the caller's initializer
            create(klass, **kwargs)._init_weight(desc, t)
            return
        if name.endswith("weight"):
            self._init_weight(name, t)
        elif name.endswith("bias"):
            self._init_bias(name, t)
        elif name.endswith("gamma"):
            self._init_gamma(name, t)
        elif name.endswith("beta"):
            self._init_beta(name, t)
        elif name.endswith("moving_mean") or name.endswith("running_mean"):
            t.zero_()
        elif name.endswith("moving_var") or name.endswith("running_var"):
            t.fill_(1.0)
        elif name.endswith("moving_inv_var") or name.endswith("moving_avg"):
            t.zero_()
        else:
            self._init_default(name, t)
        self._verbose_print(name, type(self).__name__, t)

    def _init_bias(self, _, t): t.zero_()
    def _init_gamma(self, _, t): t.fill_(1.0)
    def _init_beta(self, _, t): t.zero_()
    def _init_weight(self, name, t): raise NotImplementedError("Must override it")
    def _init_default(self, name, t): self._init_weight(name, t)


@register
class Zero(Initializer):
    def _init_weight(self, _, t): t.zero_()
    _init_default = _init_weight


@register
class One(Initializer):
    def _init_weight(self, _, t): t.fill_(1.0)
    _init_default = _init_weight


@register
class Constant(Initializer):
    def __init__(self, value=0.0):
        super().__init__(value=value); self.value = value

    def _init_weight(self, _, t):
        if hasattr(self.value, "_t"):
            t.copy_(self.value._t)
        else:
            t.fill_(float(self.value))
    _init_default = _init_weight
    _init_bias = _init_weight


@register
class Uniform(Initializer):
    def __init__(self, scale=0.07):
        super().__init__(scale=scale); self.scale = scale

    def _init_weight(self, _, t): t.uniform_(-self.scale, self.scale)


@register
class Normal(Initializer):
    def __init__(self, sigma=0.01):
        super().__init__(sigma=sigma); self.sigma = sigma

    def _init_weight(self, _, t): t.normal_(0, self.sigma)


@register
class Orthogonal(Initializer):
    def __init__(self, scale=1.414, rand_type="uniform"):
        super().__init__(scale=scale, rand_type=rand_type); self.scale, self.rand_type = scale, rand_type

    def _init_weight(self, _, t):
        nout = t.shape[0]; nin = int(np.prod(t.shape[1:]))
        tmp = torch.empty(nout, nin).uniform_(-1, 1) if self.rand_type == "uniform" else torch.randn(nout, nin)
        u, _, v = torch.linalg.svd(tmp, full_matrices=False)
        q = u if u.shape == tmp.shape else v
        t.copy_((self.scale * q).reshape(t.shape))


@register
class Xavier(Initializer):
    """``python/mxnet/initializer.py`` Xavier: scale = sqrt(magnitude / factor), fan computed with
    hw_scale = prod(shape[2:])."""

    def __init__(self, rnd_type="uniform", factor_type="avg", magnitude=3):
        super().__init__(rnd_type=rnd_type, factor_type=factor_type, magnitude=magnitude)
        self.rnd_type, self.factor_type, self.magnitude = rnd_type, factor_type, float(magnitude)

    def _init_weight(self, name, t):
        shape = t.shape
        if len(shape) < 2:
            raise ValueError("Xavier initializer cannot be applied to vector %s. It requires at least 2D." % name)
        hw_scale = float(np.prod(shape[2:])) if len(shape) > 2 else 1.0
        fan_in, fan_out = shape[1] * hw_scale, shape[0] * hw_scale
        factor = {"avg": (fan_in + fan_out) / 2.0, "in": fan_in, "out": fan_out}.get(self.factor_type)
        if factor is None:
            raise ValueError("Incorrect factor type")
        scale = math.sqrt(self.magnitude / factor)
        if self.rnd_type == "uniform":
            t.uniform_(-scale, scale)
        elif self.rnd_type == "gaussian":
            t.normal_(0, scale)
        else:
            raise ValueError("Unknown random type")


@register
class MSRAPrelu(Xavier):
    def __init__(self, factor_type="avg", slope=0.25):
        magnitude = 2.0 / (1 + slope ** 2)
        super().__init__("gaussian", factor_type, magnitude)
        self._kwargs = {"factor_type": factor_type, "slope": slope}


@register
class Bilinear(Initializer):
    def _init_weight(self, _, t):
        shape = t.shape; w = np.zeros(int(np.prod(shape)), dtype="float32")
        f = np.ceil(shape[3] / 2.0); c = (2 * f - 1 - f % 2) / (2.0 * f)
        for i in range(w.size):
            x = i % shape[3]; y = (i // shape[3]) % shape[2]
            w[i] = (1 - abs(x / f - c)) * (1 - abs(y / f - c))
        t.copy_(torch.from_numpy(w.reshape(shape)))


@register
class LSTMBias(Initializer):
    """Bias of an LSTM: zeros, except the forget gate (second quarter in MXNet's i, f, g, o order) which is ``forget_bias``
    (initializer.py LSTMBias :640-670)."""

    def __init__(self, forget_bias=1.0):
        super().__init__(forget_bias=forget_bias); self.forget_bias = forget_bias

    def _init_weight(self, _, t):
        t.zero_()
        h = t.shape[0] // 4
        t[h:2 * h] = self.forget_bias
    _init_bias = _init_default = _init_weight


@register
class FusedRNN(Initializer):
    """Initialise the flat parameter vector of a fused RNN layer: weight blocks with ``init``, biases zero, LSTM forget-gate biases
    ``forget_bias`` (initializer.py FusedRNN :673-730).  Layout: per layer and direction ``W_i2h, W_h2h`` for all, then ``b_i2h, b_h2h``."""

    def __init__(self, init, num_hidden, num_layers, mode, bidirectional=False, forget_bias=1.0):
        if isinstance(init, str):
            import json
            klass, kwargs = json.loads(init)
            init = create(klass, **kwargs)
        super().__init__(init=init.dumps() if init is not None else None, num_hidden=num_hidden, num_layers=num_layers, mode=mode,
                         bidirectional=bidirectional, forget_bias=forget_bias)
        self._init, self._h, self._layers, self._mode, self._bi, self._fb = init, num_hidden, num_layers, mode, bidirectional, forget_bias

    def _init_weight(self, desc, t):
        gates = {"rnn_relu": 1, "rnn_tanh": 1, "lstm": 4, "gru": 3}[self._mode]
        dirs, h = (2 if self._bi else 1), self._h
        flat = t.view(-1)
        nbias = self._layers * dirs * 2 * gates * h
        nweight = flat.numel() - nbias
        # input size of layer 0 follows from the total length: nweight = dirs*gates*h*(in + h) + (layers-1)*dirs*gates*h*(dirs*h + h)
        rest = (self._layers - 1) * dirs * gates * h * (dirs * h + h)
        in0 = (nweight - rest) // (dirs * gates * h) - h
        pos = 0
        for layer in range(self._layers):
            cin = in0 if layer == 0 else dirs * h
            for _ in range(dirs):
                for cols in (cin, h):
                    n = gates * h * cols
                    block = flat[pos:pos + n].view(gates * h, cols)
                    (self._init or getattr(desc, "global_init", None) or Uniform(0.07))._init_weight(str(desc).replace("parameters", "weight"), block)
                    pos += n
        bias = flat[pos:]
        bias.zero_()
        if self._mode == "lstm":
            for k in range(self._layers * dirs * 2):
                bias[k * 4 * h + h: k * 4 * h + 2 * h] = self._fb
    _init_default = _init_weight


class Mixed:
    def __init__(self, patterns, initializers):
        assert len(patterns) == len(initializers)
        self.map = list(zip([re.compile(p) for p in patterns], initializers))

    def __call__(self, name, arr):
        for prog, init in self.map:
            if prog.match(str(name)):
                init(name, arr); return
        raise ValueError("Parameter name %s did not match any pattern." % name)


class Load:
    def __init__(self, param, default_init=None, verbose=False):
        from . import ndarray as nd
        if isinstance(param, str):
            param = nd.load(param)
        self.param = {(k[4:] if k.startswith(("arg:", "aux:")) else k): v for k, v in param.items()}
        self.default_init, self.verbose = default_init, verbose

    def __call__(self, name, arr):
        if str(name) in self.param:
            src = self.param[str(name)]
            assert tuple(arr.shape) == tuple(src.shape), "Parameter %s shape mismatch" % name
            arr[:] = src
        else:
            assert self.default_init is not None, "Cannot Initialize %s" % name
            self.default_init(name, arr)
