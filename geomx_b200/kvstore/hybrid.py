"""``KVStoreHybrid`` — HiPS across several NVSwitch boxes: the box is the party, the wide-area tier runs over the native TCP stack.

This is the deployment GeoMX exists for (parties in different data centres, `docs/source/overview.rst` of the reference) mapped onto B200
boxes: inside a box the local-PS tier is a collective over NVLink (one process per GPU, ``torch.distributed``), between boxes only ONE process
per box — the box leader, local rank 0 — talks to the global server(s) with the reference's worker protocol (``kvstore/dist_ps.py`` →
``csrc/hips``).  Servers therefore see "workers" = boxes, so every server-side feature keeps working unchanged across boxes: dist_sync /
dist_async, server-side optimizers, MultiGPS, Bi-Sparse / 2-bit / fp16 transports, P3, DGT, TSEngine, HFA.

Roles / environment: start every rank of a box with ``torchrun`` (``RANK`` / ``WORLD_SIZE`` = ranks of THIS box) plus the reference's worker
environment of the box (``DMLC_ROLE=worker``, ``DMLC_PS_ROOT_URI/PORT``, ``DMLC_NUM_WORKER`` = number of boxes in the party-of-boxes ...);
servers and schedulers are started as usual.  ``rank`` / ``num_workers`` are box-local, ``num_all_workers`` = boxes x box size.

Per key and round:  box all-reduce (sum) of the pushed gradients  →  leader pushes the box aggregate  →  leader pulls the fresh value  →
box broadcast.  Collectives are NCCL on CUDA tensors and gloo on CPU tensors (the latter is what the CPU tests exercise)."""
from __future__ import annotations

import os

import torch

from ..base import MXNetError, getenv_int
from ..ndarray import NDArray
from .base import KVStoreBase


class KVStoreHybrid(KVStoreBase):
    def __init__(self, kv_type="dist_sync"):
        super().__init__(kv_type)
        import torch.distributed as dist
        self._dist = dist
        self._box_rank, self._box_size = getenv_int("RANK", 0), getenv_int("WORLD_SIZE", 1)
        if self._box_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if torch.cuda.is_available():
                local = getenv_int("LOCAL_RANK", 0)
                torch.cuda.set_device(local)
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            else:
                dist.init_process_group("gloo")
        self._leader = self._box_rank == 0
        self._tcp = None
        if self._leader:
            from .dist_ps import KVStoreDist
            self._tcp = KVStoreDist(kv_type)          # the box's single endpoint on the TCP plane
        # box-wide facts every rank needs (the leader knows them from the TCP rendezvous)
        info = [self._tcp.rank if self._leader else 0, self._tcp.num_workers if self._leader else 0,
                int(self._tcp.is_master_worker) if self._leader else 0]
        info = self._bcast_obj(info)
        self._box_index, self._num_boxes, self._master = int(info[0]), int(info[1]), bool(info[2])
        self._shapes = {}

    # ---- identity: the box is the party
    @property
    def rank(self): return self._box_rank
    @property
    def num_workers(self): return self._box_size
    @property
    def num_all_workers(self): return self._box_size * max(1, self._num_boxes)
    @property
    def is_master_worker(self): return self._master
    @property
    def box_index(self): return self._box_index
    @property
    def configures_servers(self): return self._leader and self._box_index == 0

    def _bcast_obj(self, obj):
        if self._box_size == 1:
            return obj
        box = [obj]
        self._dist.broadcast_object_list(box, src=0)
        return box[0]

    # ---- configuration goes through the leader's TCP endpoint (only box 0's leader configures the servers, like a rank-0 worker)
    def set_optimizer(self, optimizer):
        self._optimizer = optimizer
        if self._leader:
            self._tcp.set_optimizer(optimizer)

    def _set_gradient_compression(self, params):
        if self._leader:
            self._tcp.set_gradient_compression(params)

    # ---- data
    def _key(self, k):
        return k

    def _init(self, key, value):
        self._shapes[key] = tuple(value.shape)
        if self._leader:
            self._tcp.init(key, value)

    def _push(self, key, vals, priority):
        t = vals[0]._t.detach().clone()
        for v in vals[1:]:
            t.add_(v._t.detach().to(t.device))
        if self._box_size > 1:
            self._dist.all_reduce(t)                  # local-PS tier of this box: NVLink (NCCL) / gloo sum
        if self._leader:
            self._tcp.push(key, NDArray(t), priority=priority)

    def _pull(self, key, outs, priority):
        ref = outs[0]._t
        buf = torch.empty_like(ref.detach())
        if self._leader:
            tmp = NDArray(torch.empty(ref.shape, dtype=ref.dtype))
            self._tcp.pull(key, out=tmp, priority=priority)
            self._tcp.flush()
            buf.copy_(tmp._t)
        if self._box_size > 1:
            self._dist.broadcast(buf, src=0)
        for o in outs:
            tgt = o._data
            (tgt.detach() if tgt.requires_grad else tgt).copy_(buf.to(tgt.device))

    def flush(self):
        if self._tcp is not None:
            self._tcp.flush()

    def _barrier(self):
        if self._leader:
            self._tcp._barrier()
        if self._box_size > 1:
            self._dist.barrier()

    def close(self):
        if self._box_size > 1:
            self._dist.barrier()
        if self._tcp is not None:
            self._tcp.close()
        if self._box_size > 1 and self._dist.is_initialized():
            self._dist.destroy_process_group()

    def get_num_dead_node(self, node_id=7, timeout=60):
        n = self._tcp.get_num_dead_node(node_id, timeout) if self._leader else 0
        return int(self._bcast_obj(n))
