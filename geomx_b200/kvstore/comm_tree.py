"""``CommDeviceTree`` — topology-aware tree reduce / broadcast between the GPUs of ONE process.

Parity: ``src/kvstore/comm_tree.h:51-530`` (``ReduceInner`` :99-225 walks the tree bottom-up, ``BroadcastInner`` :256-329 top-down, big arrays
are split so that every GPU roots one slice: ``MXNET_KVSTORE_TREE_ARRAY_BOUND``) and ``src/kvstore/gpu_topology.h`` (link matrix
``GetP2PWeight`` :134-199, Kernighan-Lin bisection :269-400, ``ComputeTrees`` :1054).  Selected with ``MXNET_KVSTORE_USETREE=1`` for
``kv.create('device')`` exactly like the reference (``kvstore_local.h:76-86``).

The trees come from the native solver (``csrc/runtime/gpu_topology.h``); the link matrix is probed through NVML (NVLink lane count per pair)
with a uniform-weight fallback, which is also the truth on an NVSwitch box.  One tree per root is cached."""
from __future__ import annotations

import os

import torch

from ..base import getenv_int
from .local import CommDevice


def link_matrix(devices):
    """W[i][j] = relative link strength between CUDA devices (reference weights: NVLink lanes + 1 if peer access, self = 0)."""
    n = len(devices)
    W = [[0.0] * n for _ in range(n)]
    forced = os.environ.get("GEOMX_LINK_MATRIX")                     # tests / exotic hosts: "w00,w01,...;w10,..."
    if forced:
        rows = [[float(x) for x in r.split(",")] for r in forced.split(";")]
        return [row[:n] for row in rows[:n]]
    nv = None
    try:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml
    except Exception:
        nv = None
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            w = 1.0
            try:
                if devices[i].type == "cuda" and devices[j].type == "cuda" and torch.cuda.can_device_access_peer(devices[i].index, devices[j].index):
                    w += 1.0
            except Exception:
                pass
            if nv is not None and devices[i].type == "cuda":
                try:
                    hi = nv.nvmlDeviceGetHandleByIndex(devices[i].index)
                    pj = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(devices[j].index)).busId
                    for link in range(18):
                        try:
                            if nv.nvmlDeviceGetNvLinkState(hi, link) and nv.nvmlDeviceGetNvLinkRemotePciInfo(hi, link).busId == pj:
                                w += 1.0
                        except Exception:
                            break
                except Exception:
                    pass
            W[i][j] = w
    return W


def compute_tree(W, root):
    """(parent[], round[], depth) of the reduction tree rooted at ``root`` (native solver, python fallback for a missing runtime)."""
    n = len(W)
    flat = [float(W[i][j]) for i in range(n) for j in range(n)]
    from .. import runtime
    if runtime.available():
        parent, rnd, depth = runtime.C().topology_tree(flat, n, int(root))
        return list(parent), list(rnd), int(depth)
    # fallback: plain binomial tree in index order
    order = [root] + [d for d in range(n) if d != root]
    parent, rnd = [-1] * n, [-1] * n
    depth, m = 0, 1
    while m < n:
        depth += 1; m <<= 1
    step, lvl = 1, 0
    while step < n:
        for i in range(0, n, 2 * step):
            if i + step < n:
                parent[order[i + step]] = order[i]; rnd[order[i + step]] = lvl
        step <<= 1; lvl += 1
    return parent, rnd, depth


class CommDeviceTree(CommDevice):
    def __init__(self):
        super().__init__()
        self._trees = {}
        self._devices = None
        self._W = None
        self.array_bound = getenv_int("MXNET_KVSTORE_TREE_ARRAY_BOUND", 10000000)

    def _setup(self, vals):
        devs = [v._t.device for v in vals]
        if self._devices != devs:
            self._devices, self._W, self._trees = devs, link_matrix(devs), {}

    def tree(self, root):
        if root not in self._trees:
            self._trees[root] = compute_tree(self._W, root)
        return self._trees[root]

    def _reduce_tree(self, parts, root):
        """parts[i] lives on device i; returns the sum on device ``root``.  Edges of one round are independent (different senders)."""
        parent, rnd, depth = self.tree(root)
        acc = list(parts)
        for level in range(depth):
            for d, (p, r) in enumerate(zip(parent, rnd)):
                if p >= 0 and r == level:
                    acc[p] = acc[p] + acc[d].to(acc[p].device, non_blocking=True)
        return acc[root]

    def reduce(self, key, vals):
        if len(vals) == 1:
            return vals[0]._t.detach()
        self._setup(vals)
        parts = [v._t.detach() for v in vals]
        n = len(parts)
        home = self.home.get(key, parts[0].device)
        root = next((i for i, p in enumerate(parts) if p.device == home), 0)
        buf = self.merge[key]
        if parts[0].numel() < self.array_bound or n == 1:
            buf.copy_(self._reduce_tree(parts, root))
            return buf
        # big array: slice i is reduced over the tree rooted at device i (reduce-scatter), then gathered on the home device
        flat = [p.reshape(-1) for p in parts]
        bounds = [(flat[0].numel() * i) // n for i in range(n + 1)]
        out = buf.view(-1)
        for i in range(n):
            lo, hi = bounds[i], bounds[i + 1]
            if hi > lo:
                out[lo:hi].copy_(self._reduce_tree([f[lo:hi] for f in flat], i), non_blocking=True)
        return buf

    def broadcast(self, key, src, outs):
        if len(outs) <= 1 or self._devices is None or len(outs) != len(self._devices):
            return super().broadcast(key, src, outs)
        root = next((i for i, o in enumerate(outs) if o._t.device == src.device), 0)
        parent, rnd, depth = self.tree(root)
        have = {root: src}
        tgt0 = outs[root]._t
        (tgt0.detach() if tgt0.requires_grad else tgt0).copy_(src, non_blocking=True)
        for level in reversed(range(depth)):                     # top-down: the edge that fired last in the reduction goes first
            for d, (p, r) in enumerate(zip(parent, rnd)):
                if p >= 0 and r == level and p in have:
                    tgt = outs[d]._t
                    (tgt.detach() if tgt.requires_grad else tgt).copy_(have[p], non_blocking=True)
                    have[d] = tgt.detach()
