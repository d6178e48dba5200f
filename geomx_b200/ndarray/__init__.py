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


def Custom(*inputs, **kwargs):
    """``mx.nd.Custom`` — run a registered ``mx.operator.CustomOp`` (see geomx_b200/operator.py)."""
    from ..operator import Custom as _custom
    return _custom(*inputs, **kwargs)
