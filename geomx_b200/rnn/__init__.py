"""``mx.rnn`` — the legacy symbolic RNN API (parity: python/mxnet/rnn): cells that build unrolled ``mx.sym`` graphs (``rnn_cell``), checkpoint
helpers that store fused / unfused weights interchangeably (``rnn``) and the bucketing data pipeline (``io``).  New code should prefer
``mx.gluon.rnn`` (same cells, imperative)."""
from . import io, rnn, rnn_cell  # noqa: F401
from .io import BucketSentenceIter, encode_sentences  # noqa: F401
from .rnn import do_rnn_checkpoint, load_rnn_checkpoint, save_rnn_checkpoint  # noqa: F401
from .rnn_cell import *  # noqa: F401,F403
