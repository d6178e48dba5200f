"""Sentence bucketing (parity: python/mxnet/rnn/io.py:30-200)."""
from __future__ import annotations

import bisect
import random

import numpy as np

from .. import ndarray as nd
from ..io import DataBatch, DataDesc, DataIter

__all__ = ["encode_sentences", "BucketSentenceIter"]


def encode_sentences(sentences, vocab=None, invalid_label=-1, invalid_key="\n", start_label=0, unknown_token=None):
    """Token lists → integer id lists.  With ``vocab=None`` a new vocabulary is built (ids from ``start_label``, skipping
    ``invalid_label``); with a given vocabulary unknown words raise unless ``unknown_token`` is set.  Returns ``(encoded, vocab)``."""
    new_vocab = vocab is None
    if new_vocab:
        vocab = {invalid_key: invalid_label}
    idx = start_label
    res = []
    for sent in sentences:
        coded = []
        for word in sent:
            if word not in vocab:
                assert new_vocab or unknown_token, "Unknown token %s" % word
                if new_vocab:
                    if idx == invalid_label:
                        idx += 1
                    vocab[word] = idx; idx += 1
                else:
                    word = unknown_token
            coded.append(vocab[word])
        res.append(coded)
    return res, vocab


class BucketSentenceIter(DataIter):
    """Pads every sentence to the smallest bucket that holds it and serves batches of ONE bucket at a time with ``bucket_key`` set; the label
    is the data shifted left by one step (next-token prediction).  Sentences longer than the largest bucket are dropped."""

    def __init__(self, sentences, batch_size, buckets=None, invalid_label=-1, data_name="data", label_name="softmax_label", dtype="float32", layout="NT"):
        super().__init__(batch_size)
        if not buckets:
            counts = np.bincount([len(s) for s in sentences])
            buckets = [i for i, c in enumerate(counts) if c >= batch_size]
        buckets = sorted(buckets)
        self.data = [[] for _ in buckets]
        ndiscard = 0
        for s in sentences:
            b = bisect.bisect_left(buckets, len(s))
            if b == len(buckets):
                ndiscard += 1
                continue
            row = np.full((buckets[b],), invalid_label, dtype=dtype)
            row[:len(s)] = s
            self.data[b].append(row)
        self.data = [np.asarray(d, dtype=dtype) for d in self.data]
        self.ndiscard = ndiscard
        self.batch_size, self.buckets, self.invalid_label, self.dtype = batch_size, buckets, invalid_label, dtype
        self.data_name, self.label_name, self.layout = data_name, label_name, layout
        self.major_axis = layout.find("N")
        self.default_bucket_key = max(buckets)
        shape = (batch_size, self.default_bucket_key) if self.major_axis == 0 else (self.default_bucket_key, batch_size)
        self.provide_data = [DataDesc(data_name, shape, dtype, layout)]
        self.provide_label = [DataDesc(label_name, shape, dtype, layout)]
        self.idx = [(i, j) for i, d in enumerate(self.data) for j in range(0, len(d) - batch_size + 1, batch_size)]
        self.curr_idx = 0
        self.reset()

    def reset(self):
        self.curr_idx = 0
        random.shuffle(self.idx)
        for d in self.data:
            np.random.shuffle(d)
        self.nddata, self.ndlabel = [], []
        for d in self.data:
            lab = np.full_like(d, self.invalid_label)
            if d.size:
                lab[:, :-1] = d[:, 1:]
            self.nddata.append(d); self.ndlabel.append(lab)

    def next(self):
        if self.curr_idx == len(self.idx):
            raise StopIteration
        i, j = self.idx[self.curr_idx]
        self.curr_idx += 1
        d, l = self.nddata[i][j:j + self.batch_size], self.ndlabel[i][j:j + self.batch_size]
        if self.major_axis == 1:
            d, l = d.T, l.T
        data, label = nd.array(d, dtype=self.dtype), nd.array(l, dtype=self.dtype)
        return DataBatch([data], [label], pad=0, bucket_key=self.buckets[i],
                         provide_data=[DataDesc(self.data_name, tuple(data.shape), self.dtype, self.layout)],
                         provide_label=[DataDesc(self.label_name, tuple(label.shape), self.dtype, self.layout)])
