"""Shared pieces of the demo training scripts (argument parsing, the demo CNN, data sharding, evaluation).

The scripts mirror the reference's ``examples/cnn*.py`` command lines (``-lr -bs -ds -ep -ms -dc -sc -c -bcr``) and training loops, written
against ``geomx_b200`` (``import geomx_b200 as mx``).  They run in three launch styles:
  * reference style: one process per role with the ``DMLC_*`` environment (``scripts/cpu/*.sh``, ``scripts/gpu/*.sh``) over the native TCP HiPS;
  * fabric style   : ``torchrun --nproc-per-node N examples/cnn.py`` — ranks are workers, servers live in HBM shards (NVSwitch data plane);
  * stand-alone    : ``python examples/cnn.py`` (1 worker, both PS tiers collapsed).
MNIST idx files are used when present under ``--data-dir``; otherwise a deterministic synthetic set of the same shape (no network here).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402
from geomx_b200.gluon import data as gdata  # noqa: E402


def make_parser(extra=()):
    p = argparse.ArgumentParser()
    p.add_argument("-lr", "--learning-rate", type=float, default=0.01)
    p.add_argument("-bs", "--batch-size", type=int, default=32)
    p.add_argument("-ds", "--data-slice-idx", type=int, default=-1, help="-1: use the global worker rank")
    p.add_argument("-ep", "--epoch", type=int, default=5)
    p.add_argument("-sc", "--split-by-class", action="store_true")
    p.add_argument("-c", "--cpu", action="store_true")
    p.add_argument("--data-dir", default=os.environ.get("GEOMX_DATA_DIR", "/root/data"))
    p.add_argument("--max-iters", type=int, default=int(os.environ.get("GEOMX_MAX_ITERS", "0")), help="stop after this many iterations (0 = all)")
    p.add_argument("--eval-every", type=int, default=int(os.environ.get("GEOMX_EVAL_EVERY", "1")))
    p.add_argument("--load-params", default=os.environ.get("GEOMX_LOAD_PARAMS", ""), help="resume: .params file loaded into the net before kv.init "
                   "(the master worker re-initialises the global server from it)")
    p.add_argument("--save-params", default=os.environ.get("GEOMX_SAVE_PARAMS", ""), help="write the final parameters here (first training worker only)")
    if "mixed" in extra:
        p.add_argument("-ms", "--mixed-sync", action="store_true")
        p.add_argument("-dc", "--dcasgd", action="store_true")
    if "bcr" in extra:
        p.add_argument("-bcr", "--bisparse-compression-ratio", type=float, default=0.01)
    return p


def pick_context(force_cpu):
    if force_cpu:
        return mx.cpu()
    try:
        n = mx.context.num_gpus()
        ctx = mx.gpu(int(os.environ.get("LOCAL_RANK", 0)) % max(1, n))      # more worker processes than GPUs: share the devices round-robin
        mx.nd.zeros((1,), ctx=ctx)
        return ctx
    except mx.base.MXNetError:
        return mx.cpu()


def build_net(ctx, batch_size):
    if os.environ.get("GEOMX_SEED"):                    # reproducible initial weights (the reference never seeds; tests do)
        mx.random.seed(int(os.environ["GEOMX_SEED"]))
    net = mx.models.build_cnn()
    net.initialize(force_reinit=True, ctx=ctx, init=mx.init.Xavier())
    net(mx.nd.random.uniform(shape=(batch_size, 1, 28, 28), ctx=ctx))     # materialise deferred shapes
    return net


def checkpointing(net, args, kv=None):
    """Resume / save hooks shared by the demo scripts (the reference has none: "Resume in HiPS = master worker re-inits the global server from
    loaded params via kv.init").  Call right after ``build_net`` (``kv`` may be given later through the returned ``bind``): loads
    ``--load-params`` now, and writes ``--save-params`` at interpreter exit on the first training worker."""
    import atexit
    if args.load_params:
        net.load_parameters(args.load_params)
    state = {"kv": kv}

    def save():
        k = state["kv"]
        if not args.save_params or k is None or k.is_master_worker:
            return
        if worker_slice(args, k) == 0:                     # the worker that trains on data slice 0 (scripts pass -ds <global index>)
            mx.nd.waitall()
            net.save_parameters(args.save_params)
    atexit.register(save)
    return lambda k: state.__setitem__("kv", k)


def configures_servers(kv):
    """True on the process that plays the master worker: the dedicated one (reference topology) or world rank 0 (fabric / stand-alone)."""
    return kv.is_master_worker or getattr(kv, "configures_servers", False)


class _Slice(gdata.sampler.Sampler):
    def __init__(self, indices):
        self._idx = list(indices)

    def __iter__(self):
        return iter(self._idx)

    def __len__(self):
        return len(self._idx)


def make_loaders(batch_size, num_parts, part, root, by_class=False):
    root = os.path.join(os.path.expanduser(root), "mnist")
    train, test = gdata.vision.MNIST(root=root, train=True), gdata.vision.MNIST(root=root, train=False)
    n = len(train) // max(1, num_parts)
    if num_parts > 1 and by_class:
        order = sorted(range(len(train)), key=lambda i: int(train._label[i]))
        idx = order[part * n:(part + 1) * n]
    else:
        idx = range(part * n, (part + 1) * n)
    tf = gdata.vision.transforms.Compose([gdata.vision.transforms.Resize((28, 28)), gdata.vision.transforms.ToTensor()])
    tr = gdata.DataLoader(train.transform_first(tf), batch_size, sampler=_Slice(idx), last_batch="discard")
    te = gdata.DataLoader(test.transform_first(tf), batch_size, last_batch="discard")
    return tr, te


def accuracy(loader, net, ctx, max_batches=20):
    hit = tot = 0
    for b, (X, y) in enumerate(loader):
        if b >= max_batches:
            break
        X, y = X.as_in_context(ctx), y.as_in_context(ctx)
        pred = net(X).argmax(axis=1)
        hit += float((pred == y.astype("float32")).sum().asscalar()); tot += y.shape[0]
    return hit / max(1, tot)


def worker_slice(args, kv):
    if args.data_slice_idx >= 0:
        return args.data_slice_idx
    return int(os.environ.get("RANK", kv.rank))


class Progress:
    def __init__(self):
        self.t0 = time.time(); self.it = 0

    def log(self, epoch, acc):
        print("[Time %.3f][Epoch %d][Iteration %d] Test Acc %.4f" % (time.time() - self.t0, epoch, self.it, acc), flush=True)
