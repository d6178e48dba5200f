"""Token ↔ index mapping (parity: python/mxnet/contrib/text/vocab.py:30-218).

Index 0 is the unknown token, then the reserved tokens, then counter keys by descending frequency (ties: alphabetical) subject to
``most_freq_count`` / ``min_freq``."""
from __future__ import annotations

__all__ = ["Vocabulary"]


class Vocabulary:
    def __init__(self, counter=None, most_freq_count=None, min_freq=1, unknown_token="<unk>", reserved_tokens=None):
        assert min_freq > 0, "`min_freq` must be set to a positive value."
        if reserved_tokens is not None:
            rs = set(reserved_tokens)
            assert unknown_token not in rs, "`reserved_token` cannot contain `unknown_token`."
            assert len(rs) == len(reserved_tokens), "`reserved_tokens` cannot contain duplicate reserved tokens."
        self._unknown_token = unknown_token
        self._reserved_tokens = list(reserved_tokens) if reserved_tokens else None
        self._idx_to_token = [unknown_token] + (self._reserved_tokens or [])
        self._token_to_idx = {t: i for i, t in enumerate(self._idx_to_token)}
        if counter is not None:
            special = set(self._idx_to_token)
            ranked = sorted(counter.items(), key=lambda kv: kv[0])
            ranked.sort(key=lambda kv: kv[1], reverse=True)
            budget = len(ranked) if most_freq_count is None else most_freq_count
            for tok, freq in ranked:
                if freq < min_freq or budget <= 0:
                    break
                if tok not in special:
                    self._token_to_idx[tok] = len(self._idx_to_token)
                    self._idx_to_token.append(tok)
                    budget -= 1

    def __len__(self):
        return len(self._idx_to_token)

    token_to_idx = property(lambda self: self._token_to_idx)
    idx_to_token = property(lambda self: self._idx_to_token)
    unknown_token = property(lambda self: self._unknown_token)
    reserved_tokens = property(lambda self: self._reserved_tokens)

    def to_indices(self, tokens):
        """A token → its index; a list of tokens → list of indices; unknown tokens → 0."""
        if isinstance(tokens, (list, tuple)):
            return [self._token_to_idx.get(t, 0) for t in tokens]
        return self._token_to_idx.get(tokens, 0)

    def to_tokens(self, indices):
        single = not isinstance(indices, (list, tuple))
        idx = [indices] if single else indices
        out = []
        for i in idx:
            if not isinstance(i, int) or i < 0 or i >= len(self._idx_to_token):
                raise ValueError("Token index %s in the provided `indices` is invalid." % (i,))
            out.append(self._idx_to_token[i])
        return out[0] if single else out
