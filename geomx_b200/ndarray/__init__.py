"""``mx.nd`` namespace."""
from . import random  # noqa: F401
from .ndarray import *  # noqa: F401,F403
from .ndarray import NDArray  # noqa: F401
from .utils import load, load_bytes, save, save_async, save_bytes  # noqa: F401
from . import sparse  # noqa: F401,E402
from .sparse import CSRNDArray, RowSparseNDArray  # noqa: F401,E402
from . import op_lib as _op_lib  # noqa: E402
for _n in _op_lib.__all__:
    if _n not in globals():
        globals()[_n] = getattr(_op_lib, _n)
del _n
from . import contrib  # noqa: F401,E402
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
