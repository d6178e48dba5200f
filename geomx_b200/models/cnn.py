"""The reference's demo model (``examples/cnn.py:56-64``) and its B200 training engine.

``build_cnn()`` returns the Gluon network exactly as the reference scripts declare it:
``Conv2D(16,k5,relu) → MaxPool2 → Conv2D(32,k5,relu) → MaxPool2 → Dense(256,relu) → Dense(128,relu) → Dense(10)`` (178 762 parameters).

``HipsCNNTrainStep`` is the flagship hot path: one *training step* = forward + backward of that network for a per-worker batch
plus the HiPS push/pull of all 10 keys (``dist_sync``: party aggregation → global aggregation + optimizer → broadcast), captured ONCE
into a CUDA graph and replayed per step.  At the reference's batch size (32 per worker) the step is 4-5 hand-written sm_100a launches:

  cnn_fwd      conv0+bias+ReLU+pool → conv1+bias+ReLU+pool, fp32 FMA, the conv0 map never leaves shared memory      (csrc/kernels/cnn_direct.cu)
  mlp_chain    dense0 → dense1 → classifier → softmax-CE forward AND backward in one 16-CTA cluster                  (csrc/kernels/mlp_chain.cu)
  cnn_bwd_all  conv1 data gradient + pool/ReLU routing + conv0 weight gradient ‖ conv1 weight gradient, one grid      (csrc/kernels/cnn_direct.cu)
  exchange     one LL kernel for all keys after the backward pass (several ranks), or key-group channels: the dense keys' exchange on
               a high-priority stream underneath cnn_bwd_all, the conv keys' one-hop exchange after it (one rank / GEOMX_STEP_OVERLAP=1)
                                                                                                                      (csrc/kernels/hips_fabric.cu)

Larger batches (and ``GEOMX_DIRECT_CONV=0`` / ``GEOMX_FUSED_MLP=0``) use the generic kernels: direct conv0 + im2col, tcgen05 GEMMs with fused
bias / ReLU / max-pool / mask / column-sum epilogues (3xTF32 = fp32-accurate by default), weight-gradient GEMMs on a parallel graph branch.
With one rank and one party both PS tiers collapse into the arena optimizer inside the exchange kernels.

The reference runs the same step as ~200 engine ops (per-image im2col+SGEMM) + 20 ZMQ round trips.
The public API a user calls is ``step(X_host, y_host) -> loss`` (H2D of the batch from pinned memory, graph replay, D2H of the loss).
"""
from __future__ import annotations

import os

import torch

from .. import gluon, initializer
from ..context import Context
from ..ndarray import NDArray
from ..ops import native
from ..parallel.arena import ArenaLayout
from ..parallel.fabric import HipsFabric, Topology

__all__ = ["build_cnn", "CNN_PARAM_SHAPES", "HipsCNNTrainStep"]

CNN_PARAM_SHAPES = [(16, 1, 5, 5), (16,), (32, 16, 5, 5), (32,), (256, 512), (256,), (128, 256), (128,), (10, 128), (10,)]


def build_cnn():
    net = gluon.nn.Sequential()
    net.add(gluon.nn.Conv2D(channels=16, kernel_size=5, activation="relu"), gluon.nn.MaxPool2D(pool_size=2, strides=2),
            gluon.nn.Conv2D(channels=32, kernel_size=5, activation="relu"), gluon.nn.MaxPool2D(pool_size=2, strides=2),
            gluon.nn.Dense(256, activation="relu"), gluon.nn.Dense(128, activation="relu"), gluon.nn.Dense(10))
    return net


class LossHandle:
    """Result of ``HipsCNNTrainStep.step_async``: the per-sample losses of one step in pinned host memory + the event that guards them."""

    __slots__ = ("_host", "_event")

    def __init__(self, host, event):
        self._host, self._event = host, event

    def wait(self):
        if self._event is not None:
            self._event.synchronize()
        return self._host

    def item(self):
        if self._host is None:          # look-ahead engine, first call: no step has finished yet
            return float("nan")
        self._event.synchronize()
        return float(self._host.sum()) / self._host.numel()

    asscalar = item


class HipsCNNTrainStep:
    """Graph-captured training step of the demo CNN on the HiPS fabric.

    Parameters
    ----------
    net : Gluon network from :func:`build_cnn` (initialised; its parameters are re-homed into the symmetric arena so
          ``net(x)`` / ``save_parameters`` keep working on the live weights) or ``None`` (Xavier init here).
    batch_size : per-worker batch (reference default 32).
    optimizer : ``mx.optimizer.Optimizer`` with a native spec (Adam / SGD / DCASGD) — runs on the global-PS shard.
    topo : :class:`Topology` (default: from ``WORLD_SIZE`` / ``RANK`` / ``GEOMX_NUM_PARTIES`` / ``DMLC_NUM_GLOBAL_SERVER``).
    """

    def __init__(self, net=None, batch_size=32, optimizer=None, topo=None, device=None, use_graph=True,
                 use_multicast=True, mode="dist_sync", fused_zero_grad=True, wire_dtype="fp32", dgt=False, dgt_rerank_every=32,
                 update="server", bsc_threshold=None, size_lower_bound=None, hfa=None, loopback=False, lookahead=False):
        """``update='server'`` (examples/cnn.py): the optimizer runs on the global-PS shard inside the exchange kernel.  ``update='local'``
        (examples/cnn_bsc.py / cnn_fp16.py / cnn_mpq.py): the kvstore only aggregates gradients (``set_optimizer`` is not called on it), every
        worker applies its own Adam to the pulled aggregate — one fused arena-optimizer launch.  ``bsc_threshold``: Bi-Sparse between the tiers
        for keys >= ``size_lower_bound`` elements.  ``hfa=(K1, K2)`` (examples/cnn_hfa.py): local Adam on the own gradient every step, party
        average of the weights every K1 steps, global average every K1*K2 steps (reference milestone algebra, kvstore_dist_server.h:959-972).

        ``lookahead=True`` software-pipelines consecutive steps: the launch made by ``step(X[k+1], y[k+1])`` starts at the classifier head of
        batch k (whose convolutions already ran), exchanges the dense keys underneath the convolution backward pass, the conv-key exchange AND
        the forward convolutions of batch k+1, and returns the loss of batch k (``nan`` on the first call; ``flush()`` trains the last batch).
        The arithmetic and its order are unchanged — batch k+1's convolutions still read the conv weights updated by batch k, its head the
        dense weights updated by batch k — only the place where the step is cut into launches moves, so that the two-hop exchange of the
        660 KB dense keys has ~25 us of independent work to hide behind instead of ~9."""
        native.require()
        from .. import optimizer as opt
        self.B = B = int(batch_size)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.topo = topo or Topology.from_env()
        self.mode = mode
        self.hfa = tuple(int(v) for v in hfa) if hfa else None
        self.update = "local" if (update == "local" or self.hfa) else "server"
        self.fused_zero_grad = fused_zero_grad   # False keeps the gradients readable after a step (tests / debugging): memset instead
        optimizer = optimizer or opt.Adam(learning_rate=0.01)
        spec = optimizer.spec()
        if spec is None:
            raise ValueError("optimizer %s has no native spec; use Adam / SGD / DCASGD" % type(optimizer).__name__)
        self.layout = ArenaLayout.build(list(enumerate(CNN_PARAM_SHAPES)))
        self._local_spec = spec if self.update == "local" else None
        self.fabric = HipsFabric(self.layout, self.topo, self.device, None if self.update == "local" else spec, use_multicast=use_multicast,
                                 loopback=loopback)
        # the script-level `grad / num_samples`, folded into the push kernel; with local updates the aggregate is also averaged over the workers
        self.fabric.set_push_scale(1.0 / B if self.update == "server" else 1.0 / (B * self.topo.world))
        from ..base import getenv_int
        lower = int(size_lower_bound) if size_lower_bound is not None else getenv_int("MXNET_KVSTORE_SIZE_LOWER_BOUND", 200000)
        fmts = {}
        if wire_dtype in ("fp16", "mpq", "fp8") and self.fabric.protocol == "ll":
            # FP16: every key as halves on the wire (examples/cnn_fp16.py); MPQ: only the keys above MXNET_KVSTORE_SIZE_LOWER_BOUND (cnn_mpq.py:52)
            bound = lower if wire_dtype == "mpq" else 0
            fmts = {i: ("fp8" if wire_dtype == "fp8" else "fp16") for i, sl in enumerate(self.layout.slots) if sl.numel >= bound}
        if bsc_threshold and self.fabric.protocol == "ll":
            fmts.update({i: "bsc" for i, sl in enumerate(self.layout.slots) if sl.numel >= lower})
        if fmts:
            self.fabric.set_wire_formats(fmts, bsc_threshold=float(bsc_threshold or 0.01))
        self._dgt_every = 0
        if dgt and self.fabric.protocol == "ll":
            self.fabric.enable_dgt()
            self._dgt_every = max(1, int(dgt_rerank_every))
        f = self.fabric
        # Key groups ("channels") of the step: the dense keys' gradients are complete long before the convolution backward pass ends, so their
        # exchange runs underneath it on its own stream (two-hop, optimizer sharded over all ranks); the conv keys come last and use the
        # one-hop replicated mode — the only communication left on the critical path (reference ordering: push(idx, priority=-idx),
        # examples/cnn.py:121-125).  GEOMX_STEP_OVERLAP=0 selects the single fused exchange at the end of the step, =1 the channels.
        # Measured on B200 (profiles/bench_history.md): one rank -> channels (their exchange is local arithmetic); 2 ranks -> a tie; 4 and 8 ranks
        # -> the single LL exchange after the backward pass wins (0.0631 vs 0.0662 ms at 4, 0.0653 vs 0.0764 ms at 8: at 8 ranks the overlapped
        # two-hop channel takes 26-36 us underneath a 12 us backward pass and its polling slows that pass down).  Hence the default by world size.
        want_overlap = os.environ.get("GEOMX_STEP_OVERLAP", "1" if self.topo.world == 1 else "0") == "1"
        self.overlap = mode == "dist_sync" and not self.hfa and want_overlap and (self.topo.world == 1 or f.ll_d is not None)
        if self.overlap:
            f.add_channel("dense", [4, 5, 6, 7, 8, 9], replicate=False, grid=int(os.environ.get("GEOMX_DENSE_CHANNEL_GRID", 40)) or None)
            f.add_channel("conv", [0, 1, 2, 3], replicate=True)
        # GEOMX_STEP_OVERLAP=0: ONE exchange of all keys after the backward pass.  GEOMX_STEP_EXCHANGE picks its protocol: `ll` (three-hop
        # hierarchy walk), `sharded` / `replicated` (direct protocol, two hops / one hop)
        self.single = None
        one = os.environ.get("GEOMX_STEP_EXCHANGE", "ll")
        if not self.overlap and mode == "dist_sync" and not self.hfa and one in ("sharded", "replicated") and self.topo.world > 1 and f.ll_d is not None:
            f.add_channel("all", list(range(10)), replicate=one == "replicated")
            self.single = "all"
        self.fused_mlp = B <= 32 and os.environ.get("GEOMX_FUSED_MLP", "1") == "1"
        # small-batch regime: both convolutions forward / backward as direct fp32-FMA kernels (3 launches, csrc/kernels/cnn_direct.cu) instead
        # of im2col + tcgen05 GEMMs (GEOMX_DIRECT_CONV=0 selects the GEMM path; batches > 64 always use it)
        self.direct_conv = B <= 64 and B % 2 == 0 and os.environ.get("GEOMX_DIRECT_CONV", "1") == "1"
        self.P = [f.param_view(i) for i in range(10)]
        self.G = [f.grad_view(i) for i in range(10)]
        if self.update == "local":
            # worker-local weights + optimizer state: the fabric's parameter arena only receives the aggregated gradient (or, with HFA, the
            # averaged weights) in this mode
            n_el = self.layout.total
            self.Wl = torch.zeros(n_el, dtype=torch.float32, device=self.device)
            self.s0l, self.s1l = torch.zeros_like(self.Wl), torch.zeros_like(self.Wl)
            self.opt_state = torch.zeros(4, dtype=torch.int32, device=self.device)
            self.P = [self.layout.view(self.Wl, i) for i in range(10)]
        self._init_params(net)
        dev, f32 = self.device, torch.float32
        e = lambda *s, dt=f32: torch.empty(*s, dtype=dt, device=dev)
        # the batch lives in ONE device buffer [x | label] so that the per-step staging -> compute hand-over is a single D2D copy
        self.xin = e(B * 28 * 28 + B)
        self.x, self.label = self.xin[:B * 784].view(B, 1, 28, 28), self.xin[B * 784:]
        self.a1, self.idx1 = e(B, 16, 12, 12), e(B, 16, 12, 12, dt=torch.uint8)
        self.col1 = e(B * 64, 400)
        self.z2 = e(B, 32, 8, 8)
        self.a2, self.idx2 = e(B, 32, 4, 4), e(B, 32, 4, 4, dt=torch.uint8)
        self.a3, self.a4 = e(B, 256), e(B, 128)
        self.loss, self.logits = e(B), e(B, 10)
        self.dz4, self.dz3, self.da2 = e(B, 128), e(B, 256), e(B, 512)
        self.dz2rows, self.dcol1, self.da1 = e(B * 64, 32), e(B * 64, 400), e(B, 16, 12, 12)
        self.loss_host = torch.empty(B, dtype=f32).pin_memory()
        self.lookahead = bool(lookahead) and mode == "dist_sync" and not self.hfa
        # look-ahead: the head of the launch belongs to the PREVIOUS batch, whose labels must outlive the arrival of the next batch in `xin`
        self.label_cur = torch.empty_like(self.label) if self.lookahead else self.label
        self.x_cur = torch.empty_like(self.x) if self.lookahead else self.x        # ... and so must its images (conv0's weight gradient reads them)
        self._primed = False
        self.fused_exchange = False
        self.graph = None
        self._graph_alt = None
        self._xbufs = [self.xin, None]
        self._side = torch.cuda.Stream(device=self.device)
        # the exchange branch gets a high-priority stream: its (few) CTAs must become resident at once on every rank — they poll each other —
        # instead of queueing behind the convolution-backward CTAs that are launched at the same time
        self._comm = torch.cuda.Stream(device=self.device, priority=-1)
        self.use_graph = use_graph
        self.steps_done = 0
        self.kernels_per_step = 0

    # ------------------------------------------------------------------------------------------------------------
    def _init_params(self, net):
        f, topo = self.fabric, self.topo
        if net is not None:
            params = list(net.collect_params().values())
            assert [tuple(p.shape) for p in params] == CNN_PARAM_SHAPES, "network does not match the demo CNN"
            for i, p in enumerate(params):
                f.param_view(i).copy_(p.data()._t.detach().to(self.device))
        else:
            init = initializer.Xavier()
            for i, shape in enumerate(CNN_PARAM_SHAPES):
                host = torch.zeros(shape)
                init(initializer.InitDesc("w%d_%s" % (i, "weight" if len(shape) > 1 else "bias")), host)
                f.param_view(i).copy_(host.to(self.device))
        if topo.world > 1:
            import torch.distributed as dist
            dist.broadcast(f.param.tensor, src=0)      # `init`: rank-0 (master worker) value wins, then barrier
            torch.cuda.synchronize()
            dist.barrier()
        f.load_master_from_param()
        if self.update == "local":
            self.Wl.copy_(f.param.tensor)
        if net is not None:  # re-home the Gluon parameters onto the live arena (zero-copy pull)
            for i, p in enumerate(net.collect_params().values()):
                d = p.data()
                d._data = self.P[i]
                d._ctx_hint = Context("gpu", self.device.index or 0)
                if p.grad_req != "null":
                    d.attach_grad(p.grad_req)
        self.net = net

    # ------------------------------------------------------------------------------------------------------------
    def _steps(self):
        """The training step as an ordered list of (name, stream, launch); `stream` is 'main', 'side' (parallel graph branch for weight
        gradients), 'comm' (a key group's exchange on the high-priority stream) or 'join' (main, after the side branch)."""
        n = native
        B, P, G, f = self.B, self.P, self.G, self.fabric
        a2f = self.a2.view(B, 512)
        Wc1, Gc1 = P[2].view(32, 400), G[2].view(32, 400)
        kv = (lambda: (f.async_step(), f.grad.tensor.zero_() if self.fused_zero_grad else None)) if self.mode == "dist_async" else \
            (lambda: f.fsa_step(zero_grad=self.fused_zero_grad))
        if self.overlap:
            kv = lambda: f.channel_step("conv", zero_grad=self.fused_zero_grad)
        elif self.single:
            kv = lambda: f.channel_step(self.single, zero_grad=self.fused_zero_grad)
        kv_dense = [("hips push+opt+pull (dense keys, overlapped)", "comm", lambda: f.channel_step("dense", zero_grad=self.fused_zero_grad))] if self.overlap else []
        self._tail = []
        if self.update == "local":
            sp = self._local_spec
            clip = sp.get("clip_gradient", -1.0)

            def local_opt(grad_tensor, zero=None):
                n.arena_opt(sp["name"], self.Wl, grad_tensor, self.s0l, self.s1l, self.layout.total, f.tile_mult, sp["lr"], sp["wd"], sp["rescale_grad"],
                            -1.0 if clip is None else clip, sp.get("momentum", 0.0), sp.get("beta1", 0.9), sp.get("beta2", 0.999), sp.get("epsilon", 1e-8),
                            sp.get("lamda", 0.04), self.opt_state, zero)
            if self.hfa:
                # HFA: no exchange inside the step; the optimizer consumes (and clears) the worker's own gradient arena
                kv_dense = []
                kv = lambda: local_opt(f.grad.tensor, f.grad.tensor if self.fused_zero_grad else None)
            else:
                # the exchange delivered the aggregated gradient into the fabric's parameter arena: one fused arena-optimizer launch applies it
                self._tail = [("local optimizer (fused arena Adam)", "main", lambda: local_opt(f.param.tensor))]
        def conv1_fwd():
            # bias + ReLU + 2x2 max-pool in the tcgen05 epilogue (only the pooled map and its arg-max leave the SM); the two-kernel form is
            # the fallback for geometries the in-warp pooling cannot express
            if os.environ.get("GEOMX_NO_POOL_FUSION", "0") == "1" or not n.gemm_pool(self.col1, Wc1, self.a2, self.idx2, 8, 8, bias=P[3], relu=True):
                n.gemm(self.col1, Wc1, self.z2, bias=P[3], relu=True, store_nchw_hw=64)
                n.maxpool2x2_fwd(self.z2, self.a2, self.idx2)

        head = [
            ("dense0 gemm", "main", lambda: n.gemm(a2f, P[4], self.a3, bias=P[5], relu=True)),
            ("dense1 gemm", "main", lambda: n.gemm(self.a3, P[6], self.a4, bias=P[7], relu=True)),
            ("head fwd+bwd", "main", lambda: n.head_fwd_bwd(self.a4, P[8], P[9], self.label_cur, self.loss, self.logits, G[8], G[9], self.dz4, G[7], True)),
            ("dW1 gemm", "side", lambda: n.gemm(self.dz4, self.a3, G[6], a_mn=True, b_mn=True)),
            ("dz3 gemm (mask,colsum)", "main", lambda: n.gemm(self.dz4, P[6], self.dz3, b_mn=True, mask=self.a3, colsum=G[5])),
            ("dW0 gemm", "side", lambda: n.gemm(self.dz3, a2f, G[4], a_mn=True, b_mn=True)),
            ("da2 gemm", "main", lambda: n.gemm(self.dz3, P[4], self.da2, b_mn=True)),
        ]
        if self.fused_mlp:
            # dense0 -> dense1 -> classifier -> softmax-CE forward and backward: ONE 8-CTA cluster launch instead of the seven above
            head = [("mlp chain fwd+bwd (cluster)", "main", lambda: n.mlp_chain(a2f, P[4], P[5], P[6], P[7], P[8], P[9], self.label_cur, self.loss, self.logits,
                                                                              G[4], G[5], G[6], G[7], G[8], G[9], self.da2))]
        # number of leading steps that only need the batch and the conv weights (the part a look-ahead launch runs for the NEXT batch)
        self._n_fwd = 1 if self.direct_conv else 2
        carry = (self.label, self.label_cur) if self.lookahead else None
        if self.direct_conv:
            da2v = self.da2
            if os.environ.get("GEOMX_CNN_BWD_ONE_LAUNCH", "1") == "1" and B % 2 == 0 and B <= 64:
                # both backward jobs in one heterogeneous grid: no side-stream fork / join around the convolution backward
                fused = f.channel_fused_args("conv", zero_grad=self.fused_zero_grad) if self.overlap else None
                self.fused_exchange = fused is not None
                if fused is not None:
                    # ... and the conv keys' exchange in the tail of the same launch (the CTAs that finish last serve one tile each)
                    blk, tiles, n_act = fused
                    return [
                        ("conv0+conv1 fwd (direct, relu+pool fused)", "main", lambda: n.cnn_fwd(self.x, P[0], P[1], P[2], P[3], self.a1, self.idx1, self.a2, self.idx2, carry=carry, x_keep=self.x_cur if self.lookahead else None)),
                    ] + head + kv_dense + [
                        ("conv bwd (one grid) + conv-key exchange in its tail", "main",
                         lambda: n.cnn_bwd_exchange(self.x_cur, P[2], self.a1, self.idx1, self.a2, self.idx2, da2v, G[0], G[1], G[2], G[3], blk, tiles, n_act)),
                    ] + self._tail
                return [
                    ("conv0+conv1 fwd (direct, relu+pool fused)", "main", lambda: n.cnn_fwd(self.x, P[0], P[1], P[2], P[3], self.a1, self.idx1, self.a2, self.idx2, carry=carry, x_keep=self.x_cur if self.lookahead else None)),
                ] + head + kv_dense + [
                    ("conv bwd: conv1 wgrad + dgrad + conv0 wgrad (direct, one grid)", "main",
                     lambda: n.cnn_bwd_all(self.x_cur, P[2], self.a1, self.idx1, self.a2, self.idx2, da2v, G[0], G[1], G[2], G[3])),
                    ("hips push+opt+pull", "join", kv),
                ] + self._tail
            return [
                ("conv0+conv1 fwd (direct, relu+pool fused)", "main", lambda: n.cnn_fwd(self.x, P[0], P[1], P[2], P[3], self.a1, self.idx1, self.a2, self.idx2, carry=carry, x_keep=self.x_cur if self.lookahead else None)),
            ] + head + kv_dense + [
                ("conv1 wgrad (direct, sparse)", "side", lambda: n.cnn_wgrad1(self.a1, self.a2, self.idx2, da2v, G[2], G[3])),
                ("conv1 dgrad + conv0 wgrad (direct)", "main", lambda: n.cnn_bwd(self.x_cur, P[2], self.a1, self.idx1, self.a2, self.idx2, da2v, G[0], G[1])),
                ("hips push+opt+pull", "join", kv),
            ] + self._tail
        if self.lookahead:
            conv1 = conv1_fwd
            conv1_fwd = lambda: (conv1(), self.label_cur.copy_(self.label), self.x_cur.copy_(self.x))
        return [
            ("conv0+relu+pool+im2col", "main", lambda: n.conv_relu_pool_im2col_fwd(self.x, P[0], P[1], self.a1, self.idx1, self.col1, 5, 5)),
            ("conv1 gemm (bias,relu,maxpool fused)", "main", conv1_fwd),
        ] + head + kv_dense + [
            ("pool+relu bwd -> rows", "main", lambda: n.pool_relu_bwd_rows(self.da2.view(B, 32, 4, 4), self.a2, self.idx2, self.dz2rows, G[3])),
            ("dWc1 gemm (split-K)", "side", lambda: n.gemm(self.dz2rows, self.col1, Gc1, a_mn=True, b_mn=True, split_k=16, accumulate=True)),
            ("dcol1 gemm", "main", lambda: n.gemm(self.dz2rows, Wc1, self.dcol1, b_mn=True)),
            ("conv0 wgrad + col2im", "main", lambda: n.conv_relu_pool_wgrad_col2im(self.x_cur, self.dcol1, self.a1, self.idx1, G[0], G[1], CNN_PARAM_SHAPES[0], 5, 5)),
            ("hips push+opt+pull", "join", kv),
        ] + self._tail

    def _body(self, stop_after=None, part="all"):
        """Issue the step (``part='all'``), only its forward convolutions (``'fwd'``), everything after them (``'rest'``), or the look-ahead
        rotation ``rest(batch k) + fwd(batch k+1)`` (``'rotated'``)."""
        before = native.launch_count
        f = self.fabric
        # the gradient arena is cleared by the previous step's HiPS kernel (fused zero_grad) unless gradients must stay readable
        if not self.fused_zero_grad:
            f.grad.tensor.zero_()
        main, side, comm = torch.cuda.current_stream(), self._side, self._comm
        forked = comm_forked = False
        steps = self._steps()
        nf = self._n_fwd
        steps = {"all": steps, "fwd": steps[:nf], "rest": steps[nf:], "rotated": steps[nf:] + steps[:nf]}[part]
        for i, (name, where, fn) in enumerate(steps):
            if stop_after is not None and i >= stop_after:
                break
            if where == "side":      # weight-gradient GEMMs leave the critical path: parallel branch of the captured graph
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    fn()
                forked = True
            elif where == "comm":    # a key group's exchange: third branch, after everything issued so far on main and side
                comm.wait_stream(main)
                if forked:           # (a stream that has no work of this step yet must not be waited on: it is not part of a graph capture)
                    comm.wait_stream(side)
                with torch.cuda.stream(comm):
                    fn()
                comm_forked = True
            else:
                if where == "join" and forked:
                    main.wait_stream(side); forked = False
                fn()
        if forked:
            main.wait_stream(side)
        if comm_forked:
            main.wait_stream(comm)
        if part in ("all", "rotated"):
            self.kernels_per_step = native.launch_count - before

    def capture(self):
        """Warm up (2 eager steps on a side stream) and capture the step into a CUDA graph."""
        if self.graph is not None or not self.use_graph:
            return
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
                self.steps_done += 1
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self._dgt_every:
            self.fabric.dgt_rerank()      # first ranking now: the collective's lazy communicator set-up belongs to the warm-up, not to a timed step
        if self.topo.world > 1:
            import torch.distributed as dist
            dist.barrier()
        if self.lookahead:
            self._prime()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body(part="rotated" if self.lookahead else "all")
        self.graph = g
        self._graph_alt = None

    def _prime(self):
        """Look-ahead: run the forward convolutions of the batch in ``self.x`` so that the next launch can start at the classifier head."""
        if not self._primed:
            self._body(part="fwd")
            self._primed = True

    def flush(self):
        """Look-ahead: train on the batch whose convolutions are already done (the last one enqueued) and return its mean loss."""
        if not (self.lookahead and self._primed):
            return None
        self._body(part="rest")
        self._primed = False
        self.steps_done += 1
        return float(self.loss.mean())

    def _bind_input(self, k):
        """Point the step at input buffer ``k`` (0 = ``self.xin``; 1 = the second staging buffer of the host pipeline).  Only matters while
        launches are being issued or captured: a captured graph keeps the buffer it was captured with."""
        buf = self._xbufs[k]
        B = self.B
        self.xin, self.x, self.label = buf, buf[:B * 784].view(B, 1, 28, 28), buf[B * 784:]
        if not self.lookahead:
            self.x_cur, self.label_cur = self.x, self.label

    def _alt_graph(self):
        """The same step captured on input buffer 1: the host pipeline alternates between the two graphs, so a batch is consumed where the
        H2D copy put it (no staging -> compute copy on the critical path of the step)."""
        if self._graph_alt is None:
            self._bind_input(1)
            try:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(s):
                    with torch.cuda.graph(g, stream=s):
                        self._body()
                torch.cuda.current_stream().wait_stream(s)
            finally:
                self._bind_input(0)
            self._graph_alt = g
        return self._graph_alt

    def run_device(self, buf=0):
        """One step on whatever is currently in input buffer ``buf`` (0: ``self.x`` / ``self.label``; device-only, used by the kernel-time
        bench and by the host pipeline)."""
        if self._dgt_every and self.steps_done and self.steps_done % self._dgt_every == 0 and self.graph is not None:
            self.fabric.dgt_rerank()          # outside the captured graph; order / formats are updated in place
        if self.use_graph:
            if self.graph is None:
                self.capture()
            if self.lookahead:
                self._prime()
            (self._alt_graph() if buf == 1 else self.graph).replay()
        elif buf == 1:
            self._bind_input(1)
            try:
                self._body()
            finally:
                self._bind_input(0)
        elif self.lookahead:
            self._prime()
            self._body(part="rotated")
        else:
            self._body()
        self.steps_done += 1
        if self.hfa and self.steps_done % self.hfa[0] == 0:
            self._hfa_sync(global_round=(self.steps_done // self.hfa[0]) % self.hfa[1] == 0)

    def _hfa_sync(self, global_round):
        """Average the workers' weights inside the party (every K1 steps) or over the whole job (every K1*K2 steps).  Outside the captured
        step: the party round is ONE launch of the party all-reduce kernel, the global round one launch of the fused exchange without an
        optimizer (push_scale 1/parties turns the sum of party means into their mean)."""
        f, topo = self.fabric, self.topo
        if topo.world == 1:
            return
        f.grad.tensor.copy_(self.Wl).mul_(1.0 / topo.party_size)
        if global_round:
            scale = f.push_scale
            f.set_push_scale(1.0 / topo.num_parties)
            f.fsa_step()
            f.set_push_scale(scale)
        else:
            f.party_allreduce(f.grad, f.param, scale=1.0)
        self.Wl.copy_(f.param.tensor)
        f.grad.tensor.zero_()

    def _pipeline(self):
        if not hasattr(self, "_pl"):
            dev, B = self.device, self.B
            # direct mode: the two H2D targets ARE the step's two input buffers (one captured graph each); look-ahead steps and DGT runs
            # keep the single-graph form with a staging -> compute copy
            self._direct_inputs = not self.lookahead and os.environ.get("GEOMX_E2E_DIRECT_INPUT", "1") == "1"
            if self._direct_inputs:
                self._xbufs[1] = torch.empty_like(self.xin)
                stage = [self._xbufs[0], self._xbufs[1]]
            else:
                stage = [torch.empty_like(self.xin) for _ in range(2)]
            self._pl = {
                "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev), "stage": stage,
                "sx": [s[:B * 784].view(B, 1, 28, 28) for s in stage], "sy": [s[B * 784:] for s in stage],
                "ready": [torch.cuda.Event() for _ in range(2)], "done": [torch.cuda.Event() for _ in range(2)],
                "ring": [(torch.empty_like(self.loss, device="cpu").pin_memory(), torch.cuda.Event()) for _ in range(4)], "n": 0, "last_loss": None,
            }
        return self._pl

    def step_async(self, X, y):
        """Public API: enqueue one training step and return a :class:`LossHandle` without waiting for the GPU (the MXNet engine's contract:
        an op returns at once, ``asscalar()`` synchronises).  ``X`` (B,1,28,28) and ``y`` (B,) are host tensors / NDArrays (pinned → async
        H2D) or device tensors.

        Three streams: the batch is copied host→device on a copy stream into one of the step's TWO input buffers (overlapping the previous
        step's compute); the compute stream replays the graph that was captured on that buffer (two graphs of the same step, alternating —
        no staging→compute copy; ``GEOMX_E2E_DIRECT_INPUT=0`` and look-ahead steps use one graph plus a device-to-device copy), and the
        per-sample loss goes device→host on a third stream into a pinned ring slot; ``handle.item()`` waits for exactly that copy.  A
        training loop therefore launches step i+1 before it reads the loss of step i."""
        X = X._t if isinstance(X, NDArray) else X
        y = y._t if isinstance(y, NDArray) else y
        pl = self._pipeline()
        i = pl["n"]; b = i % 2
        pl["n"] = i + 1
        main = torch.cuda.current_stream()
        if X.is_cuda:
            self.x.copy_(X.reshape(self.x.shape), non_blocking=True)
            self.label.copy_(y.reshape(self.label.shape), non_blocking=True)
        else:
            h2d = pl["h2d"]
            if i >= 2:
                h2d.wait_event(pl["done"][b])              # staging buffer b was handed over by step i-2
            with torch.cuda.stream(h2d):
                pl["sx"][b].copy_(X.reshape(self.x.shape), non_blocking=True)
                pl["sy"][b].copy_(y.reshape(self.label.shape), non_blocking=True)
                pl["ready"][b].record(h2d)
            main.wait_event(pl["ready"][b])
            if not self._direct_inputs:
                self.xin.copy_(pl["stage"][b], non_blocking=True)
        if pl["last_loss"] is not None:
            main.wait_event(pl["last_loss"])                # the loss buffer of the previous step has been read out
        if self.lookahead and not self._primed:
            # first batch of a look-ahead run: only its convolutions can run yet (after the graph warm-up, which also uses this batch);
            # there is no finished step to report
            if self.use_graph and self.graph is None:
                self.capture()
            self._prime()
            pl["done"][b].record(main)
            return LossHandle(None, None)
        self.run_device(buf=b if (not X.is_cuda and self._direct_inputs) else 0)
        pl["done"][b].record(main)
        host, ev = pl["ring"][i % 4]
        d2h = pl["d2h"]
        d2h.wait_event(pl["done"][b])
        with torch.cuda.stream(d2h):
            host.copy_(self.loss, non_blocking=True)
            ev.record(d2h)
        pl["last_loss"] = ev
        return LossHandle(host, ev)

    def step(self, X, y):
        """Public API: one training step; returns the mean loss as a Python float (forces the D2H read of the per-sample loss)."""
        return self.step_async(X, y).item()

    # reference semantics helpers ------------------------------------------------------------------------------------
    def params_numpy(self):
        return [p.detach().cpu().numpy().copy() for p in self.P]

    def h2d_bytes_per_step(self):
        return self.x.numel() * 4 + self.label.numel() * 4

    def d2h_bytes_per_step(self):
        return self.loss.numel() * 4
