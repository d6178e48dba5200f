"""Single-process KVStores: ``local`` / ``device`` / ``nccl``.

Parity: ``src/kvstore/kvstore_local.h:69-454`` (Init/Push/Pull grouping, updater vs. assign, str↔int key map,
row_sparse pull), ``src/kvstore/comm.h`` (CommCPU :104-438 reduce on host, CommDevice :448-792 reduce on one
GPU with P2P, optional 2-bit compressed inter-GPU reduce :545-589), ``comm_tree.h`` (topology-aware trees) and
``kvstore_nccl.h:62-551`` (rooted reduce + broadcast).

B200 design: with one process per GPU the intra-process multi-device path is the exception, not the rule, and
NVSwitch makes the topology uniform — the reference's Kernighan-Lin tree solver (``gpu_topology.h``) collapses
to "reduce on the key's home GPU, round-robin by key".  ``CommDevice`` therefore gathers through CUDA P2P
(`cudaMemcpyPeerAsync` under ``Tensor.to``) and sums with ONE n-ary add launch (``native.nary_sum``) rather
than a pairwise chain; ``CommCPU`` stages through pinned host memory.
"""
from __future__ import annotations

import torch

from ..base import MXNetError, getenv_int
from ..ndarray import NDArray
from . import compression as gc
from .base import KVStoreBase

__all__ = ["KVStoreLocal", "KVStoreNCCL", "CommCPU", "CommDevice"]


class CommCPU:
    """Reduce on host (``comm.h:104-438``)."""

    def __init__(self):
        self.merge = {}

    def init(self, key, like):
        self.merge[key] = torch.zeros_like(like._t.detach(), device="cpu")

    def reduce(self, key, vals):
        if len(vals) == 1:
            return vals[0]._t.detach()
        buf = self.merge[key]
        buf.copy_(vals[0]._t.detach())
        for v in vals[1:]:
            buf.add_(v._t.detach().to("cpu"))
        return buf

    def broadcast(self, key, src, outs):
        for o in outs:
            tgt = o._t
            (tgt.detach() if tgt.requires_grad else tgt).copy_(src, non_blocking=True)


class CommDevice(CommCPU):
    """Reduce on the key's home GPU through P2P + one n-ary sum launch (``comm.h:448-792``)."""

    def __init__(self):
        super().__init__()
        self.home = {}
        self._gc = None

    def set_gradient_compression(self, comp):
        self._gc = comp

    def init(self, key, like):
        t = like._t.detach()
        self.home[key] = t.device
        self.merge[key] = torch.zeros_like(t)

    def reduce(self, key, vals):
        if len(vals) == 1:
            return vals[0]._t.detach()
        dev = self.home.get(key, vals[0]._t.device)
        buf = self.merge[key]
        if self._gc is not None and self._gc.active and buf.dtype == torch.float32:
            # compressed inter-GPU path (comm.h:545-589): quantise on the source GPU, ship 1/16th, dequantise+sum
            buf.zero_()
            for i, v in enumerate(vals):
                q = self._gc.quantize(("dev", key, i), v._t.detach())
                buf.add_(self._gc.dequantize(q.to(dev), buf.numel()).view_as(buf))
            return buf
        parts = [v._t.detach().to(dev, non_blocking=True) for v in vals]
        from ..ops import native
        if dev.type == "cuda" and native.available() and buf.dtype == torch.float32:
            native.nary_sum(buf, parts)
        else:
            torch.sum(torch.stack(parts), dim=0, out=buf)
        return buf


class KVStoreLocal(KVStoreBase):
    def __init__(self, kv_type="local"):
        super().__init__(kv_type)
        use_dev = "device" in kv_type
        if use_dev and getenv_int("MXNET_KVSTORE_USETREE", 0):
            from .comm_tree import CommDeviceTree              # kvstore_local.h:76-86
            self._comm = CommDeviceTree()
        else:
            self._comm = CommDevice() if use_dev else CommCPU()
        self._store = {}
        self._key_type = None
        self._next_str_key = 0
        self._gc = gc.GradientCompression()

    def _key(self, k):
        kt = str if isinstance(k, str) else int
        if self._key_type is None:
            self._key_type = kt
        elif self._key_type is not kt:
            raise MXNetError("inconsistent key types: mixing str and int keys is not allowed")
        if kt is str:
            if k not in self._str_key_map:
                self._str_key_map[k] = self._next_str_key; self._next_str_key += 1
            return self._str_key_map[k]
        return int(k)

    def _init(self, key, value):
        if key in self._store:
            raise MXNetError("duplicate init of key %s" % key)
        self._store[key] = NDArray(value._t.detach().clone())
        self._comm.init(key, value)

    def _push(self, key, vals, priority):
        if key not in self._store:
            raise MXNetError("key %s has not been inited" % key)
        merged = self._comm.reduce(key, vals)
        stored = self._store[key]
        if self._updater_func is not None:
            m = merged
            if m.device != stored._t.device:
                m = m.to(stored._t.device)
            self._updater_func(key, NDArray(m), stored)
        else:
            if merged.device != stored._t.device or merged.dtype != stored._t.dtype:
                stored._t = merged.to(stored._t.device).clone()
            else:
                stored._t.copy_(merged)

    def _pull(self, key, outs, priority):
        if key not in self._store:
            raise MXNetError("key %s has not been inited" % key)
        self._comm.broadcast(key, self._store[key]._t, outs)

    def _row_sparse_pull(self, key, outs, row_ids, priority):
        """``out`` may be a RowSparseNDArray (receives exactly the unique requested rows) or a dense NDArray (other rows zeroed) —
        kvstore_local.h:357-417 PullRowSparseImpl."""
        from ..ndarray.sparse import RowSparseNDArray, _gather
        from .utils import unique_rows
        src = self._store[key]._t
        for o, ids in zip(outs, row_ids):
            rows = unique_rows(ids._t.to(src.device))
            picked = _gather(src.contiguous(), rows)
            if isinstance(o, RowSparseNDArray):
                dev = o.data._t.device
                o._set_rows(picked.to(dev), rows.to(dev))
                o._shape = tuple(src.shape)
            else:
                tgt = o._t
                tgt.zero_()
                tgt[rows.to(tgt.device)] = picked.to(tgt.device)

    def _set_gradient_compression(self, params):
        if "device" not in self._type:
            raise MXNetError("Gradient compression is not supported for this type of kvstore")
        self._gc.set_params(params)
        self._comm.set_gradient_compression(self._gc)


class KVStoreNCCL(KVStoreLocal):
    """``kv.create('nccl')`` — per-key rooted ``ncclReduce`` + updater + ``ncclBcast`` over the process's GPUs
    (``kvstore_nccl.h:213-421``).  Uses ``torch.cuda.nccl`` (single-process communicator clique)."""

    def __init__(self, kv_type="nccl"):
        super().__init__("device")
        self._type = "nccl"

    def _push(self, key, vals, priority):
        if len(vals) > 1 and all(v._t.is_cuda for v in vals):
            import torch.cuda.nccl as nccl
            ins = [v._t.detach().contiguous() for v in vals]
            root = 0
            out = torch.empty_like(ins[root])
            outs = [out if i == root else torch.empty_like(t) for i, t in enumerate(ins)]
            nccl.reduce(ins, output=outs[root], root=root)
            merged = outs[root]
            stored = self._store[key]
            if self._updater_func is not None:
                self._updater_func(key, NDArray(merged.to(stored._t.device)), stored)
            else:
                stored._t.copy_(merged)
            return
        super()._push(key, vals, priority)

    def _pull(self, key, outs, priority):
        if len(outs) > 1 and all(o._t.is_cuda for o in outs):
            import torch.cuda.nccl as nccl
            src = self._store[key]._t
            tensors = [o._t.detach() for o in outs]
            root_dev = src.device
            ridx = next((i for i, t in enumerate(tensors) if t.device == root_dev), 0)
            tensors[ridx].copy_(src)
            nccl.broadcast(tensors, root=ridx)
            return
        super()._pull(key, outs, priority)
