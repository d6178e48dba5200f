"""Token counting (parity: python/mxnet/contrib/text/utils.py:28-85)."""
from __future__ import annotations

import collections
import re

__all__ = ["count_tokens_from_str"]


def count_tokens_from_str(source_str, token_delim=" ", seq_delim="\n", to_lower=False, counter_to_update=None):
    """Counts the tokens of ``source_str``; sequences are separated by ``seq_delim``, tokens by ``token_delim`` (both are treated as
    literal strings, empty tokens are dropped).  Returns a ``collections.Counter`` (``counter_to_update`` updated in place when given)."""
    parts = re.split("%s|%s" % (re.escape(token_delim), re.escape(seq_delim)), source_str)
    toks = [t.lower() if to_lower else t for t in parts if t]
    if counter_to_update is None:
        return collections.Counter(toks)
    counter_to_update.update(toks)
    return counter_to_update
