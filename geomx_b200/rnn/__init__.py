"""``mx.rnn`` — the bucketing data pipeline of the legacy RNN API (parity: python/mxnet/rnn/io.py ``encode_sentences`` / ``BucketSentenceIter``).

The symbolic cell classes of ``python/mxnet/rnn/rnn_cell.py`` are superseded by ``mx.gluon.rnn`` (same cells, imperative); sequence models
that need per-length graphs use ``mx.mod.BucketingModule`` with this iterator."""
from .io import BucketSentenceIter, encode_sentences  # noqa: F401
