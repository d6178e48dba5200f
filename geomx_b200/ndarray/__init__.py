"""``mx.nd`` namespace."""
from . import random  # noqa: F401
from .ndarray import *  # noqa: F401,F403
from .ndarray import NDArray  # noqa: F401
from .utils import load, load_bytes, load_frombuffer, save, save_async, save_bytes  # noqa: F401
from . import sparse  # noqa: F401,E402
from .sparse import CSRNDArray, RowSparseNDArray  # noqa: F401,E402
from . import op_lib as _op_lib  # noqa: E402
for _n in _op_lib.__all__:
    if _n not in globals():
        globals()[_n] = getattr(_op_lib, _n)
del _n
from . import contrib  # noqa: F401,E402
from . import _internal, image, linalg, op  # noqa: F401,E402
from .random import (exponential as random_exponential, gamma as random_gamma, generalized_negative_binomial as random_generalized_negative_binomial,  # noqa: F401,E402
                     multinomial as sample_multinomial, negative_binomial as random_negative_binomial, normal as random_normal, poisson as random_poisson,
                     randint as random_randint, uniform as random_uniform)
sample_uniform, sample_normal, sample_gamma, sample_exponential, sample_poisson = random_uniform, random_normal, random_gamma, random_exponential, random_poisson
sample_negative_binomial, sample_generalized_negative_binomial = random_negative_binomial, random_generalized_negative_binomial
from .random import shuffle  # noqa: F401,E402


def Custom(*inputs, **kwargs):
    """``mx.nd.Custom`` — run a registered ``mx.operator.CustomOp`` (see geomx_b200/operator.py)."""
    from ..operator import Custom as _custom
    return _custom(*inputs, **kwargs)


# ---- fluent methods: ``x.sin()``, ``x.slice_axis(...)``, ``x.topk(...)`` … (ndarray.py of the reference defines one method per registered op)
def _attach_fluent():
    names = ["arccos", "arccosh", "arcsin", "arcsinh", "arctan", "arctanh", "argmax_channel", "argsort", "broadcast_axes", "broadcast_like", "cbrt",
             "ceil", "cos", "cosh", "degrees", "depth_to_space", "diag", "expm1", "fix", "flip", "floor", "log10", "log1p", "log2", "nanprod",
             "nansum", "ones_like", "pad", "pick", "prod", "radians", "rcbrt", "reciprocal", "repeat", "reshape_like", "rint", "round", "rsqrt",
             "shape_array", "sin", "sinh", "size_array", "slice", "slice_axis", "slice_like", "softmin", "sort", "space_to_depth", "split",
             "swapaxes", "take", "tan", "tile", "trunc", "zeros_like", "topk", "one_hot", "erf", "sign", "norm", "clip", "softmax",
             "log_softmax", "relu", "sigmoid", "tanh", "exp", "log", "sqrt", "square", "abs", "expand_dims", "squeeze", "flatten", "transpose"]
    g = globals()
    for n in names:
        fn = g.get(n)
        if fn is not None and not hasattr(NDArray, n):
            setattr(NDArray, n, (lambda f: lambda self, *a, **k: f(self, *a, **k))(fn))


_attach_fluent()


def concatenate(arrays, axis=0, always_copy=True):
    """Join a list of arrays along an existing axis (legacy helper; ``always_copy=False`` returns the single input itself)."""
    assert isinstance(arrays, (list, tuple)) and len(arrays) > 0
    if not always_copy and len(arrays) == 1:
        return arrays[0]
    import torch as _torch
    return NDArray(_torch.cat([a._t for a in arrays], dim=axis))


def onehot_encode(indices, out):
    """Legacy one-hot: writes ``out[i, indices[i]] = 1`` (``out`` gives depth and dtype)."""
    out._t.zero_()
    out._t.scatter_(1, indices._t.long().view(-1, 1), 1)
    return out


def true_divide(lhs, rhs):
    return divide(lhs, rhs)  # noqa: F405


def to_dlpack_for_read(data):
    """DLPack capsule sharing the array's memory (zero copy; consumers must not write)."""
    import torch.utils.dlpack as _dl
    return _dl.to_dlpack(data._t.detach())


def to_dlpack_for_write(data):
    import torch.utils.dlpack as _dl
    return _dl.to_dlpack(data._t.detach())


def from_dlpack(dlpack):
    """Wrap a DLPack capsule (or any object with ``__dlpack__``) without copying."""
    import torch.utils.dlpack as _dl
    return NDArray(_dl.from_dlpack(dlpack))


NDArray.to_dlpack_for_read = lambda self: to_dlpack_for_read(self)
NDArray.to_dlpack_for_write = lambda self: to_dlpack_for_write(self)


def imdecode(str_img, clip_rect=(0, 0, 0, 0), out=None, index=0, channels=3, mean=None):
    """Decode an encoded image buffer to an HWC uint8 NDArray (legacy ``mx.nd.imdecode``; see ``mx.image.imdecode``)."""
    from ..image import imdecode as _imdecode
    img = _imdecode(str_img, flag=1 if channels == 3 else 0)
    if clip_rect != (0, 0, 0, 0):
        x0, y0, x1, y1 = clip_rect
        img = img[y0:y1, x0:x1]
    if mean is not None:
        img = img.astype("float32") - mean
    return img
