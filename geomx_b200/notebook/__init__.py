"""``mx.notebook`` — training callbacks for interactive sessions (parity: python/mxnet/notebook/callback.py)."""
from . import callback  # noqa: F401
