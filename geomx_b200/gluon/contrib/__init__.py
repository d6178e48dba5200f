"""``mx.gluon.contrib`` — layers outside the core namespace (parity: python/mxnet/gluon/contrib)."""
from . import nn  # noqa: F401
