"""``mx.gluon.contrib`` — layers outside the core namespace (parity: python/mxnet/gluon/contrib/{nn,rnn,data})."""
from . import data, nn, rnn  # noqa: F401
