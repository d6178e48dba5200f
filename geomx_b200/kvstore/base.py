"""KVStore base class, factory and the global flush registry.

Parity: ``python/mxnet/kvstore.py:99-705`` (KVStore API: init / push / pull / row_sparse_pull /
set_gradient_compression / set_optimizer / type / rank / num_workers / num_all_workers / is_master_worker /
save_optimizer_states / load_optimizer_states / _set_updater / _barrier / _send_command_to_servers) and the
factory ``create`` (:663-705) → ``KVStore::Create`` (``src/kvstore/kvstore.cc:41-82``: type strings
``local | device | nccl | dist_sync | dist_async | dist_device_sync | dist_sync_device``).
"""
from __future__ import annotations

import weakref

from .. import profiler as _prof
from ..base import MXNetError
from ..ndarray import NDArray
from ..ndarray.sparse import RowSparseNDArray

__all__ = ["KVStoreBase", "create", "flush_all", "register_flushable"]

_flushables = weakref.WeakSet()


def register_flushable(obj):
    """Objects with a ``flush()`` method drained by ``mx.nd.waitall()``."""
    _flushables.add(obj)


def flush_all():
    for o in list(_flushables):
        o.flush()


def _ctype_key_value(keys, vals):
    """Normalise (key | [keys], value | [values] | [[values]]) → list of (key, [values])."""
    if isinstance(keys, (tuple, list)):
        assert len(keys) == len(vals), "keys and values must have the same length"
        out = []
        for k, v in zip(keys, vals):
            out.extend(_ctype_key_value(k, v))
        return out
    if isinstance(vals, (NDArray, RowSparseNDArray)):
        return [(keys, [vals])]
    for v in vals:
        assert isinstance(v, (NDArray, RowSparseNDArray)), "value must be NDArray or list of NDArray"
    return [(keys, list(vals))]


class KVStoreBase:
    """Common Python surface; subclasses implement ``_init/_push/_pull`` on normalised (key, [values])."""

    def __init__(self, kv_type):
        self._type = kv_type
        self._updater = None
        self._updater_func = None
        self._str_key_map = {}
        self._compression = {"type": "none"}
        register_flushable(self)

    # -- key handling: str and int keys are both allowed but never mixed (kvstore_local.h:276-303)
    def _key(self, k):
        return k

    def init(self, key, value):
        for k, vals in _ctype_key_value(key, value):
            self._init(self._key(k), vals[0])

    def push(self, key, value, priority=0):
        if _prof._state["running"]:
            with _prof.scope("KVStorePush", "kvstore"):                # the reference's engine-op names (kvstore_dist.h:601, comm.h:171)
                return self._push_all(key, value, priority)
        return self._push_all(key, value, priority)

    def _push_all(self, key, value, priority):
        for k, vals in _ctype_key_value(key, value):
            if any(isinstance(v, RowSparseNDArray) for v in vals):
                # row_sparse gradients (kvstore_dist.h PushRowSparse :628-657): stores with a sparse wire get the rows, the rest the dense view
                self._push_row_sparse(self._key(k), vals, priority)
            else:
                self._push(self._key(k), vals, priority)

    def _push_row_sparse(self, key, vals, priority):
        self._push(key, [v.tostype("default") if isinstance(v, RowSparseNDArray) else v for v in vals], priority)

    def pull(self, key, out=None, priority=0, ignore_sparse=True):
        assert out is not None
        if _prof._state["running"]:
            with _prof.scope("KVStorePull", "kvstore"):
                for k, outs in _ctype_key_value(key, out):
                    self._pull(self._key(k), outs, priority)
            return
        for k, outs in _ctype_key_value(key, out):
            self._pull(self._key(k), outs, priority)

    def row_sparse_pull(self, key, out=None, priority=0, row_ids=None):
        assert out is not None and row_ids is not None
        pairs = _ctype_key_value(key, out)
        if isinstance(row_ids, NDArray):
            row_ids = [row_ids] * sum(len(o) for _, o in pairs)
        elif len(pairs) > 1 and len(row_ids) == len(pairs) and any(len(o) != 1 for _, o in pairs):
            row_ids = [r for r, (_, o) in zip(row_ids, pairs) for _ in o]
        flat_ids = list(row_ids)
        i = 0
        for k, outs in pairs:
            ids = flat_ids[i:i + len(outs)]; i += len(outs)
            self._row_sparse_pull(self._key(k), outs, ids, priority)

    def set_gradient_compression(self, compression_params):
        ctype = compression_params.get("type", "none")
        if ctype not in ("none", "2bit", "bsc"):
            raise MXNetError("Unknown type for gradient compression %s" % ctype)
        self._compression = dict(compression_params)
        self._set_gradient_compression(self._compression)

    def set_optimizer(self, optimizer):
        from .. import optimizer as opt
        self._set_updater(opt.get_updater(optimizer))

    def _set_updater(self, updater):
        self._updater = updater
        self._updater_func = updater

    @property
    def type(self): return self._type
    @property
    def rank(self): return 0
    @property
    def num_workers(self): return 1
    @property
    def num_all_workers(self): return 1
    @property
    def is_master_worker(self): return False

    def save_optimizer_states(self, fname, dump_optimizer=False):
        assert self._updater is not None, "Cannot save states for distributed training"
        with open(fname, "wb") as f:
            f.write(self._updater.get_states(dump_optimizer))

    def load_optimizer_states(self, fname):
        assert self._updater is not None, "Cannot load states for distributed training"
        with open(fname, "rb") as f:
            self._updater.set_states(f.read())

    def _barrier(self): pass
    def _send_command_to_servers(self, head, body): pass
    def flush(self): pass
    def _set_gradient_compression(self, params): pass

    def _row_sparse_pull(self, key, outs, row_ids, priority):
        raise NotImplementedError

    def get_num_dead_node(self, node_id=0, timeout=60):
        return 0


def create(name="local"):
    """Create a KVStore.  ``dist*`` types start (or attach to) the HiPS runtime."""
    if not isinstance(name, str):
        raise TypeError("name must be a string")
    n = name.lower()
    if n in ("local", "local_update_cpu", "local_allreduce_cpu", "device", "local_allreduce_device"):
        from .local import KVStoreLocal
        return KVStoreLocal(name)
    if n == "nccl":
        from .local import KVStoreNCCL
        return KVStoreNCCL(name)
    if n.startswith("dist"):
        from .dist import create_dist
        return create_dist(name)
    raise MXNetError("Unknown KVStore type \"%s\"" % name)
