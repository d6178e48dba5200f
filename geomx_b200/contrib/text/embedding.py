"""Pre-trained token embeddings (parity: python/mxnet/contrib/text/embedding.py:40-705).

``register`` / ``create`` / ``get_pretrained_file_names`` form the registry; ``CustomEmbedding`` loads any ``token v1 v2 …`` text file;
``CompositeEmbedding`` concatenates several embeddings over one vocabulary.  ``GloVe`` and ``FastText`` know the published file names but
never download: the file must already be under ``embedding_root`` (default ``$MXNET_HOME/embeddings/<name>/``) — this framework is built
for air-gapped clusters."""
from __future__ import annotations

import io
import os
import warnings

import numpy as np

from ... import ndarray as nd
from . import vocab

__all__ = ["register", "create", "get_pretrained_file_names", "GloVe", "FastText", "CustomEmbedding", "CompositeEmbedding"]

_REGISTRY = {}


def register(embedding_cls):
    """Class decorator: makes ``create(<lower-case class name>)`` build the class."""
    _REGISTRY[embedding_cls.__name__.lower()] = embedding_cls
    return embedding_cls


def create(embedding_name, **kwargs):
    cls = _REGISTRY.get(embedding_name.lower())
    if cls is None:
        raise KeyError("Cannot find `embedding_name` %s. Valid names: %s" % (embedding_name, sorted(_REGISTRY)))
    return cls(**kwargs)


def get_pretrained_file_names(embedding_name=None):
    if embedding_name is not None:
        if embedding_name.lower() not in _REGISTRY:
            raise KeyError("Cannot find `embedding_name` %s. Valid names: %s" % (embedding_name, sorted(_REGISTRY)))
        return list(getattr(_REGISTRY[embedding_name.lower()], "pretrained_file_names", ()))
    return {n: list(getattr(c, "pretrained_file_names", ())) for n, c in _REGISTRY.items()}


class _TokenEmbedding(vocab.Vocabulary):
    """A vocabulary whose tokens carry vectors.  Index 0 (unknown) gets ``init_unknown_vec(shape)``."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._vec_len, self._idx_to_vec = 0, None

    @staticmethod
    def _root(embedding_root, name):
        root = embedding_root or os.path.join(os.environ.get("MXNET_HOME", os.path.join("~", ".mxnet")), "embeddings")
        return os.path.join(os.path.expanduser(root), name)

    def _load_embedding(self, path, elem_delim, init_unknown_vec, encoding="utf8"):
        if not os.path.isfile(path):
            raise ValueError("`pretrained_file_path` %s is not a file (pre-trained embeddings are never downloaded; place the file there)" % path)
        vecs, seen, vec_len, unk_vec = [], set(), None, None
        with io.open(path, "r", encoding=encoding) as f:
            for ln, line in enumerate(f, 1):
                elems = line.rstrip().split(elem_delim)
                if len(elems) <= 2:                      # fastText header line "count dim" or a 1-d vector: skip like the reference
                    if len(elems) == 2 and ln == 1:
                        continue
                    warnings.warn("line %d of %s: unexpected format, skipped" % (ln, path))
                    continue
                tok, vals = elems[0], elems[1:]
                if tok == self.unknown_token and unk_vec is None:
                    unk_vec = [float(v) for v in vals]; vec_len = vec_len or len(vals)
                    continue
                if tok in seen:
                    warnings.warn("line %d of %s: duplicate token %s, first vector kept" % (ln, path, tok))
                    continue
                if vec_len is None:
                    vec_len = len(vals)
                assert len(vals) == vec_len, "line %d of %s: vector length %d != %d" % (ln, path, len(vals), vec_len)
                seen.add(tok)
                self._token_to_idx[tok] = len(self._idx_to_token)
                self._idx_to_token.append(tok)
                vecs.append(vals)
        self._vec_len = vec_len or 0
        n_special = len(self._idx_to_token) - len(vecs)
        table = np.zeros((len(self._idx_to_token), self._vec_len), dtype=np.float32)
        if vecs:
            table[n_special:] = np.asarray(vecs, dtype=np.float32)
        table[0] = np.asarray(unk_vec, dtype=np.float32) if unk_vec is not None else init_unknown_vec(shape=(self._vec_len,)).asnumpy()
        self._idx_to_vec = nd.array(table)

    def _build_for_vocabulary(self, vocabulary, sources):
        """Re-index onto ``vocabulary``: row i = concatenation of each source's vector of token i (unknown vector when absent)."""
        self._idx_to_token = list(vocabulary.idx_to_token)
        self._token_to_idx = dict(vocabulary.token_to_idx)
        self._unknown_token, self._reserved_tokens = vocabulary.unknown_token, vocabulary.reserved_tokens
        parts = [src.get_vecs_by_tokens(self._idx_to_token).asnumpy() for src in sources]
        table = np.concatenate(parts, axis=1)
        self._vec_len, self._idx_to_vec = table.shape[1], nd.array(table)

    vec_len = property(lambda self: self._vec_len)
    idx_to_vec = property(lambda self: self._idx_to_vec)

    def get_vecs_by_tokens(self, tokens, lower_case_backup=False):
        single = not isinstance(tokens, (list, tuple))
        toks = [tokens] if single else tokens
        if lower_case_backup:
            idx = [self._token_to_idx[t] if t in self._token_to_idx else self._token_to_idx.get(t.lower(), 0) for t in toks]
        else:
            idx = [self._token_to_idx.get(t, 0) for t in toks]
        vecs = nd.Embedding(nd.array(idx), self._idx_to_vec)
        return vecs[0] if single else vecs

    def update_token_vectors(self, tokens, new_vectors):
        assert self._idx_to_vec is not None, "The property `idx_to_vec` has not been properly set."
        toks = [tokens] if not isinstance(tokens, (list, tuple)) else tokens
        nv = new_vectors.reshape((len(toks), -1)) if len(new_vectors.shape) == 1 else new_vectors
        assert nv.shape == (len(toks), self._vec_len), "The length of new_vectors must be equal to the number of tokens and the width to vec_len."
        for t in toks:
            if t not in self._token_to_idx:
                raise ValueError("Token %s is unknown. To update the embedding vector for an unknown token, please specify it explicitly as the "
                                 "`unknown_token` %s in `tokens`." % (t, self._idx_to_token[0]))
        table = self._idx_to_vec.asnumpy().copy()
        table[[self._token_to_idx[t] for t in toks]] = nv.asnumpy()
        self._idx_to_vec = nd.array(table)


def _known_file_embedding(name, files, default):
    def __init__(self, pretrained_file_name=default, embedding_root=None, init_unknown_vec=nd.zeros, vocabulary=None, **kwargs):
        if pretrained_file_name not in files:
            raise KeyError("Cannot find pretrained file %s for token embedding %s. Valid files: %s" % (pretrained_file_name, name.lower(), ", ".join(files)))
        _TokenEmbedding.__init__(self, **kwargs)
        self._load_embedding(os.path.join(self._root(embedding_root, name.lower()), pretrained_file_name), " ", init_unknown_vec)
        if vocabulary is not None:
            self._build_for_vocabulary(vocabulary, [_Snapshot(self)])
    return register(type(name, (_TokenEmbedding,), {"__init__": __init__, "pretrained_file_names": tuple(files),
                                                    "__doc__": "%s vectors from a local copy of one of %s." % (name, ", ".join(files[:3]) + ", …")}))


class _Snapshot:
    """Frozen view of an embedding, used while the same object is being re-indexed."""

    def __init__(self, emb):
        self._t2i, self._vecs = dict(emb.token_to_idx), emb.idx_to_vec

    def get_vecs_by_tokens(self, tokens):
        return nd.Embedding(nd.array([self._t2i.get(t, 0) for t in tokens]), self._vecs)


GloVe = _known_file_embedding("GloVe", ["glove.42B.300d.txt", "glove.6B.50d.txt", "glove.6B.100d.txt", "glove.6B.200d.txt", "glove.6B.300d.txt",
                                         "glove.840B.300d.txt", "glove.twitter.27B.25d.txt", "glove.twitter.27B.50d.txt",
                                         "glove.twitter.27B.100d.txt", "glove.twitter.27B.200d.txt"], "glove.840B.300d.txt")
FastText = _known_file_embedding("FastText", ["wiki.simple.vec", "wiki.en.vec", "wiki.zh.vec", "wiki.de.vec", "wiki.fr.vec", "crawl-300d-2M.vec",
                                               "wiki-news-300d-1M.vec"], "wiki.simple.vec")


@register
class CustomEmbedding(_TokenEmbedding):
    """User file: ``<token><elem_delim><v1><elem_delim><v2>…`` per line."""

    def __init__(self, pretrained_file_path, elem_delim=" ", encoding="utf8", init_unknown_vec=nd.zeros, vocabulary=None, **kwargs):
        super().__init__(**kwargs)
        self._load_embedding(pretrained_file_path, elem_delim, init_unknown_vec, encoding)
        if vocabulary is not None:
            self._build_for_vocabulary(vocabulary, [_Snapshot(self)])


@register
class CompositeEmbedding(_TokenEmbedding):
    """Vectors of several embeddings concatenated, indexed by ``vocabulary``."""

    def __init__(self, vocabulary, token_embeddings):
        embs = token_embeddings if isinstance(token_embeddings, (list, tuple)) else [token_embeddings]
        for e in embs:
            assert isinstance(e, _TokenEmbedding), "The argument `token_embeddings` must be an instance or a list of instances of token embeddings."
        vocab.Vocabulary.__init__(self)
        self._build_for_vocabulary(vocabulary, embs)
