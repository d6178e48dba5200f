"""DataLoader.  Parity: ``python/mxnet/gluon/data/dataloader.py`` (default_batchify_fn, DataLoader with sampler /
batch_sampler / last_batch / shuffle / num_workers).

B200 design: batches are assembled into **pinned host memory** (so the H2D copy of each step is a true async DMA)
and, with ``num_workers>0``, by a pool of prefetching threads (``2 * num_workers`` batches in flight; tensor stacking, numpy and image
decoding release the GIL); ``worker_type="process"`` selects the reference's model instead — worker PROCESSES that assemble batches into
shared memory (``mx.cpu_shared``; python/mxnet/gluon/data/dataloader.py:28-120 with CPUSharedStorageManager) — for datasets whose per-item
Python transform is GIL-bound.  ``num_workers=0`` (what the reference examples use) assembles inline.  Datasets that expose
``_fast_batch(indices)`` (the in-memory vision sets) skip per-item Python entirely."""
from __future__ import annotations

import numpy as np
import torch

from ...ndarray import NDArray
from . import sampler as _sampler

__all__ = ["DataLoader", "default_batchify_fn"]


_IN_WORKER = False          # worker processes never touch CUDA: no pinning there, batches go to shared memory instead
_WORKER_STATE = {}


def _worker_init(dataset, batchify_fn):
    global _IN_WORKER
    _IN_WORKER = True
    torch.set_num_threads(1)
    _WORKER_STATE["dataset"], _WORKER_STATE["batchify"] = dataset, batchify_fn


def _to_shared(obj):
    if isinstance(obj, NDArray):
        return obj._t.share_memory_()
    if isinstance(obj, (list, tuple)):
        return [_to_shared(o) for o in obj]
    return obj


def _from_shared(obj):
    if isinstance(obj, torch.Tensor):
        return NDArray(_pin(obj))
    if isinstance(obj, list):
        return [_from_shared(o) for o in obj]
    return obj


def _worker_batch(indices):
    ds, bf = _WORKER_STATE["dataset"], _WORKER_STATE["batchify"]
    return _to_shared(bf([ds[i] for i in indices]))


def _pin(t):
    if _IN_WORKER:
        return t
    if torch.cuda.is_available():
        try:
            return t.pin_memory()
        except RuntimeError:
            return t
    return t


def default_batchify_fn(data):
    if isinstance(data[0], NDArray):
        return NDArray(_pin(torch.stack([d._t for d in data])))
    if isinstance(data[0], torch.Tensor):
        return NDArray(_pin(torch.stack(data)))
    if isinstance(data[0], tuple):
        return [default_batchify_fn(list(i)) for i in zip(*data)]
    arr = np.asarray(data)
    if arr.dtype == np.float64:
        arr = arr.astype(np.float32)
    return NDArray(_pin(torch.from_numpy(arr)))


class DataLoader:
    def __init__(self, dataset, batch_size=None, shuffle=False, sampler=None, last_batch=None, batch_sampler=None,
                 batchify_fn=None, num_workers=0, pin_memory=False, prefetch=None, thread_pool=False, worker_type=None):
        self._dataset = dataset
        if batch_sampler is None:
            if batch_size is None:
                raise ValueError("batch_size must be specified unless batch_sampler is specified")
            if sampler is None:
                sampler = _sampler.RandomSampler(len(dataset)) if shuffle else _sampler.SequentialSampler(len(dataset))
            elif shuffle:
                raise ValueError("shuffle must not be specified if sampler is specified")
            batch_sampler = _sampler.BatchSampler(sampler, batch_size, last_batch if last_batch else "keep")
        elif batch_size is not None or shuffle or sampler is not None or last_batch is not None:
            raise ValueError("batch_size, shuffle, sampler and last_batch must not be specified if batch_sampler is specified.")
        self._batch_sampler = batch_sampler
        self._batchify_fn = batchify_fn or default_batchify_fn
        self._num_workers = max(0, int(num_workers))
        if worker_type not in (None, "thread", "process"):
            raise ValueError("worker_type must be 'thread' or 'process'")
        self._process_workers = worker_type == "process" and self._num_workers > 0
        self._prefetch = max(1, int(prefetch)) if prefetch else 2 * max(1, self._num_workers)
        self._fast = getattr(dataset, "_fast_batch", None)

    def __iter__(self):
        if self._fast is not None and self._batchify_fn is default_batchify_fn:
            for batch in self._batch_sampler:
                yield self._fast(batch)
            return
        if self._num_workers == 0:
            for batch in self._batch_sampler:
                yield self._batchify_fn([self._dataset[i] for i in batch])
            return
        from collections import deque
        if self._process_workers:
            import torch.multiprocessing as tmp
            # fork keeps unpicklable datasets usable, but a forked child must not inherit a live CUDA context: spawn once CUDA is up
            ctx = tmp.get_context("spawn" if torch.cuda.is_initialized() else "fork")
            with ctx.Pool(self._num_workers, initializer=_worker_init, initargs=(self._dataset, self._batchify_fn)) as pool:
                q = deque()
                it = iter(self._batch_sampler)
                for b in it:
                    q.append(pool.apply_async(_worker_batch, (list(b),)))
                    if len(q) >= self._prefetch:
                        break
                while q:
                    r = q.popleft()
                    nxt = next(it, None)
                    if nxt is not None:
                        q.append(pool.apply_async(_worker_batch, (list(nxt),)))
                    yield _from_shared(r.get())
            return
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(self._num_workers) as pool:
            q = deque()
            it = iter(self._batch_sampler)
            def submit():
                try:
                    b = next(it)
                except StopIteration:
                    return False
                q.append(pool.submit(lambda bb=b: self._batchify_fn([self._dataset[i] for i in bb])))
                return True
            for _ in range(2 * self._num_workers):
                if not submit():
                    break
            while q:
                f = q.popleft(); submit()
                yield f.result()

    def __len__(self):
        return len(self._batch_sampler)
