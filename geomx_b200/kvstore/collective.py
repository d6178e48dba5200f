"""``mx.kv.create('dist_sync' | 'dist_async')`` for ``torchrun``-launched CPU ranks: the HiPS semantics of the fabric KVStore expressed with
``torch.distributed`` (gloo) collectives.

The NVSwitch fabric (``parallel/fabric_kvstore.py``) needs GPUs; the TCP plane (``kvstore/dist_ps.py``) needs the reference's role
environment.  This third, small implementation covers the remaining case — one process per rank started by ``torchrun`` on a machine
without GPUs (CI, laptops, the CPU half of the test-suite) — with the same observable behaviour as the fabric store:

* topology from ``RANK`` / ``WORLD_SIZE`` / ``GEOMX_NUM_PARTIES`` / ``DMLC_NUM_GLOBAL_SERVER`` (``parallel.fabric.Topology``): ``rank`` and
  ``num_workers`` are party-local, ``num_all_workers`` is the world size, ``configures_servers`` is True on world rank 0;
* ``init``: world rank 0's value wins (broadcast); ``push`` accumulates, ``pull`` is lazy — the round runs when a pulled array is read or at
  ``mx.nd.waitall()``: party sum (the local tier), sum over parties (the global tier), then the optimizer — configured on rank 0 and
  replicated — updates an identical master copy on every rank (what the global server would broadcast);
* without an optimizer the aggregate itself is what workers pull (``cnn_bsc.py`` / ``cnn_hfa.py`` flows with a local ``Trainer``);
* HFA (``MXNET_KVSTORE_USE_HFA``): rounds with ``local_iters % K2 != 0`` stop after the party tier;
* ``dist_async``: every party's aggregate is applied in party order (the deterministic serialisation of MixedSync's arrival order).

It is a plumbing path: nothing here is a hot loop, and the two-tier reduction is two ``all_reduce`` calls on sub-groups.
Parity: ``python/mxnet/kvstore.py:118-394`` (API), ``src/kvstore/kvstore_dist_server.h:1213-1366`` (FSA round), ``:959-972`` (HFA)."""
from __future__ import annotations

import os

import torch

from ..base import MXNetError, getenv_int
from .base import KVStoreBase

__all__ = ["KVStoreCollective"]


class KVStoreCollective(KVStoreBase):
    def __init__(self, kv_type="dist_sync"):
        super().__init__(kv_type)
        import torch.distributed as dist
        from ..parallel.fabric import Topology
        world = getenv_int("WORLD_SIZE", 1)
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            dist.init_process_group("gloo")
        self._dist = dist if world > 1 else None
        self._topo = Topology.from_env()
        self._sync = "async" not in kv_type
        self._store, self._pending, self._pulls, self._order = {}, {}, [], []
        self._hfa = getenv_int("MXNET_KVSTORE_USE_HFA", 0) != 0
        self._hfa_k2 = max(1, getenv_int("MXNET_KVSTORE_HFA_K2", 1))
        self._local_iters = 0
        self._configured, self._optimizer = False, None
        self._party_group = self._leaders_group = None
        if self._dist is not None and self._topo.num_parties > 1:
            # every rank must create every group, in the same order
            for g in range(self._topo.num_parties):
                grp = dist.new_group(self._topo.party_ranks(g))
                if g == self._topo.party:
                    self._party_group = grp

    # -- identity ---------------------------------------------------------------------------------------------------------------
    @property
    def rank(self): return self._topo.local
    @property
    def num_workers(self): return self._topo.party_size
    @property
    def num_all_workers(self): return self._topo.world
    @property
    def is_master_worker(self): return False
    @property
    def configures_servers(self): return self._topo.rank == 0

    # -- configuration ------------------------------------------------------------------------------------------------------------
    def set_optimizer(self, optimizer):
        self._optimizer = optimizer
        from ..optimizer import get_updater
        self._updater = get_updater(optimizer)

    def _set_gradient_compression(self, params):
        # the links between CPU ranks are loopback / shared memory: the setting is recorded (and broadcast) but nothing is compressed
        pass

    def _sync_configuration(self):
        """Rank 0's optimizer / compression setting reaches every rank before the first round (kController / kSetGradientCompression)."""
        if self._configured:
            return
        self._configured = True
        if self._dist is None:
            return
        import pickle
        blob = [pickle.dumps((self._optimizer, self._compression)) if self._topo.rank == 0 else None]
        self._dist.broadcast_object_list(blob, src=0)
        if self._topo.rank != 0:
            opt, comp = pickle.loads(blob[0])
            if opt is not None:
                self.set_optimizer(opt)
            self._compression = comp

    # -- data -----------------------------------------------------------------------------------------------------------------------
    def _init(self, key, value):
        if key in self._store:
            raise MXNetError("duplicate init of key %s" % key)
        t = value._t.detach().clone().float() if value._t.dtype != torch.float32 else value._t.detach().clone()
        if self._dist is not None:
            self._dist.broadcast(t, src=0)
        self._store[key] = t
        self._order.append(key)

    def _push(self, key, vals, priority):
        if key not in self._store:
            raise MXNetError("key %s has not been initialised" % key)
        acc = vals[0]._t.detach().float().clone().reshape(self._store[key].shape)
        for v in vals[1:]:
            acc += v._t.detach().float().reshape(acc.shape)
        if key in self._pending:
            self._pending[key] += acc
        else:
            self._pending[key] = acc

    def _pull(self, key, outs, priority):
        if key not in self._store:
            raise MXNetError("key %s has not been initialised" % key)
        self._pulls.append((key, outs))
        for o in outs:
            o._pending = self.flush

    def _all_reduce(self, t, group=None):
        if self._dist is not None:
            self._dist.all_reduce(t, group=group)
        return t

    def flush(self):
        if not self._pending and not self._pulls:
            return
        pulls, self._pulls = self._pulls, []
        for _, outs in pulls:
            for o in outs:
                o._pending = None
        if self._pending:
            self._sync_configuration()
            topo = self._topo
            keys = [k for k in self._order if k in self._pending]           # identical order on every rank
            if self._dist is not None:
                n = torch.tensor([len(keys)]); lo = n.clone(); self._dist.all_reduce(lo, op=self._dist.ReduceOp.MIN)
                if int(lo) != len(keys):
                    raise MXNetError("collective KVStore: ranks pushed different key sets this round")
            local_only = False
            if self._hfa:
                self._local_iters += 1
                local_only = self._local_iters % self._hfa_k2 != 0
            for k in keys:
                g = self._pending.pop(k)
                party_sum = self._all_reduce(g, self._party_group if topo.num_parties > 1 else None)      # local tier
                if local_only:
                    self._store[k] = party_sum                       # HFA local synchronisation: the party aggregate is what workers pull
                    continue
                if self._hfa:
                    # global HFA round: mean over parties of the party values (milestone algebra with identical milestones)
                    tot = party_sum.clone()
                    if topo.num_parties > 1:
                        self._all_reduce(tot)                        # counts every party `party_size` times
                        tot /= topo.party_size
                    self._store[k] = tot / topo.num_parties
                    continue
                if self._sync or topo.num_parties == 1:
                    tot = party_sum.clone()
                    if topo.num_parties > 1:
                        self._all_reduce(tot); tot /= topo.party_size
                    self._apply(k, tot)
                else:
                    # MixedSync: party aggregates are applied one after the other, in party order on every rank
                    parts = [torch.zeros_like(party_sum) for _ in range(topo.world)]
                    self._dist.all_gather(parts, party_sum)
                    for p in range(topo.num_parties):
                        self._apply(k, parts[p * topo.party_size])
        for k, outs in pulls:
            src = self._store[k]
            for o in outs:
                tgt = o._data
                (tgt.detach() if tgt.requires_grad else tgt).copy_(src.reshape(tgt.shape).to(tgt.dtype))

    def _apply(self, key, grad):
        from ..ndarray import NDArray
        if self._updater is not None:
            idx = self._order.index(key) if not isinstance(key, int) else key
            self._updater(idx, NDArray(grad), NDArray(self._store[key]))
        else:
            self._store[key] = grad                                    # no optimizer on the "server": the aggregate is stored

    def _barrier(self):
        self.flush()
        if self._dist is not None:
            self._dist.barrier()

    def get_num_dead_node(self, node_id=0, timeout=60):
        return 0

    def save_optimizer_states(self, fname, dump_optimizer=False):
        assert self._updater is not None, "Cannot save states for distributed training without an optimizer"
        if self._topo.rank == 0:
            with open(fname, "wb") as f:
                f.write(self._updater.get_states(dump_optimizer))
        self._barrier()

    def load_optimizer_states(self, fname):
        assert self._updater is not None
        with open(fname, "rb") as f:
            self._updater.set_states(f.read())
