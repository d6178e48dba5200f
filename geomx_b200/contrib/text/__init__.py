"""``mx.contrib.text`` — vocabulary indexing and pre-trained token embeddings (parity: python/mxnet/contrib/text/{vocab,embedding,utils}.py)."""
from . import embedding, utils, vocab  # noqa: F401
from .vocab import Vocabulary  # noqa: F401
