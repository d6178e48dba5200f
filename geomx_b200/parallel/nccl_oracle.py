"""NCCL + cuDNN/cuBLAS "oracle" of the flagship step: the same schedule expressed with library ops only.

This is BASELINE.md's fallback comparator ("our NCCL+cuBLAS oracle path implementing the identical schedule") and the correctness oracle of
SURVEY §7.2-4: PyTorch conv/linear (cuDNN/cuBLAS), autograd, hierarchical gradient aggregation with NCCL (`all_reduce` inside the party, then
across party leaders, then broadcast — or one world all-reduce when there is a single party), `grad / num_samples` scaling and Adam with the
same hyper-parameters on every rank, all captured in a CUDA graph.  None of the hand-written kernels is on this path."""
from __future__ import annotations


import torch
import torch.nn.functional as F

from ..models.cnn import CNN_PARAM_SHAPES
from .. import initializer


class OracleCNNTrainStep:
    def __init__(self, batch_size=32, optimizer=None, topo=None, device=None, use_graph=True):
        self.B, self.topo, self.device, self.use_graph = batch_size, topo, device, use_graph
        self.lr = optimizer.lr if optimizer is not None else 0.01
        init = initializer.Xavier()
        self.P = []
        for i, shape in enumerate(CNN_PARAM_SHAPES):
            host = torch.zeros(shape)
            init(initializer.InitDesc("w%d_%s" % (i, "weight" if len(shape) > 1 else "bias")), host)
            self.P.append(host.to(device).requires_grad_(True))
        if topo.world > 1:
            import torch.distributed as dist
            for p in self.P:
                dist.broadcast(p.data, src=0)
        self.m = [torch.zeros_like(p) for p in self.P]
        self.v = [torch.zeros_like(p) for p in self.P]
        self.t = torch.zeros((), device=device)
        self.x = torch.empty(batch_size, 1, 28, 28, device=device)
        self.label = torch.empty(batch_size, device=device)
        self.loss = torch.zeros(batch_size, device=device)
        self.loss_host = torch.empty(batch_size).pin_memory()
        self.graph = None
        self.kernels_per_step = 0
        self.fabric = None
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True

    def _body(self):
        P = self.P
        h = F.max_pool2d(torch.relu(F.conv2d(self.x, P[0], P[1])), 2)
        h = F.max_pool2d(torch.relu(F.conv2d(h, P[2], P[3])), 2).flatten(1)
        h = torch.relu(F.linear(h, P[4], P[5])); h = torch.relu(F.linear(h, P[6], P[7]))
        loss = F.cross_entropy(F.linear(h, P[8], P[9]), self.label.long(), reduction="none")
        grads = torch.autograd.grad(loss.sum(), P)
        self.loss.copy_(loss.detach())
        flat = torch.cat([g.reshape(-1) for g in grads]) / self.B
        if self.topo.world > 1:
            import torch.distributed as dist
            dist.all_reduce(flat)
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        lr_t = self.lr * torch.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        off = 0
        with torch.no_grad():
            for p, m, v in zip(P, self.m, self.v):
                g = flat[off:off + p.numel()].view_as(p); off += p.numel()
                m.mul_(b1).add_(g, alpha=1 - b1); v.mul_(b2).addcmul_(g, g, value=1 - b2)
                p.sub_(lr_t * m / (v.sqrt() + eps))

    def capture(self):
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._body()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body()
        self.graph = g

    def run_device(self):
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._body()

    def step(self, X, y):
        self.x.copy_(X.reshape(self.x.shape), non_blocking=True); self.label.copy_(y.reshape(self.label.shape), non_blocking=True)
        self.run_device()
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host.mean())

    def h2d_bytes_per_step(self): return self.x.numel() * 4 + self.label.numel() * 4
    def d2h_bytes_per_step(self): return self.loss.numel() * 4
