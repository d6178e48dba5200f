#!/usr/bin/env python
"""Reference arm of bench.py: the UNMODIFIED INET-RC/GeoMX build (MXNet 1.4.0, installed under baseline/_ref, see baseline/README.md)
training the reference's own demo CNN (examples/cnn.py:56-64) through the reference's own public API.

Nothing of geomx_b200 is imported here.  The loop is the one examples/cnn.py:101-133 runs — forward/backward under autograd.record, then for
every parameter ``kv.push(idx, grad / num_samples, priority=-idx)`` and ``kv.pull(idx, out=param, priority=-idx)`` with Adam set on the
kvstore — with the one change the offline build forces: the kvstore is the reference's in-process multi-GPU store (``kv.create('nccl')``,
src/kvstore/kvstore_nccl.h, or ``device``) instead of ``dist_sync``, because ps-lite's ZeroMQ/protobuf transport cannot be built without
network access (USE_DIST_KVSTORE=0).  N GPUs = one process driving N devices with per-device batch 32 (weak scaling), the stock MXNet 1.4
data-parallel path; under torchrun only rank 0 works, the other ranks exit.

Timing: MXNet exposes no CUDA events, so both numbers are host-clocked between ``mx.nd.waitall()`` barriers (device idle on both sides).
``value`` = K steps on device-resident batches; ``e2e`` = K steps that each copy the batch from host memory (H2D) and read the loss (D2H).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def ensure_library():
    """baseline/_ref ships libmxnet.so xz-compressed (270 MB -> 40 MB); unpack it once next to the package."""
    so = os.path.join(REF, "mxnet", "libmxnet.so")
    xz = so + ".xz"
    if os.path.exists(so) and (not os.path.exists(xz) or os.path.getmtime(so) >= os.path.getmtime(xz)):
        return so
    if not os.path.exists(xz):
        raise FileNotFoundError("baseline/_ref/mxnet/libmxnet.so(.xz) missing — run baseline/build_reference.sh")
    import lzma
    tmp = so + ".tmp.%d" % os.getpid()
    with lzma.open(xz, "rb") as fi, open(tmp, "wb") as fo:
        while True:
            b = fi.read(1 << 24)
            if not b:
                break
            fo.write(b)
    os.chmod(tmp, 0o755)
    os.replace(tmp, so)
    return so


def reexec_with_library_path():
    """libmxnet.so needs libopenblas.so (our miniblas shim, same directory) and the CUDA 12.9 toolkit libraries on the loader path."""
    if os.environ.get("GEOMX_REF_ENV") == "1":
        return
    dirs = [os.path.join(REF, "mxnet"), "/usr/local/cuda/lib64", "/usr/lib/x86_64-linux-gnu"]
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(dirs + [env.get("LD_LIBRARY_PATH", "")])
    env["GEOMX_REF_ENV"] = "1"
    env["PYTHONPATH"] = REF + os.pathsep + env.get("PYTHONPATH", "")
    env.setdefault("MXNET_CUDNN_AUTOTUNE_DEFAULT", "0")
    env.setdefault("OMP_NUM_THREADS", "4")
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--kvstore", default=None, help="nccl | device | local (default: nccl for N>1, device for N=1)")
    ap.add_argument("--cpu", action="store_true", help="plumbing test without a GPU")
    ap.add_argument("--impl", default="reference")
    args, _ = ap.parse_known_args()
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0          # single-process multi-device job: rank 0 drives all N GPUs
    try:
        ensure_library()
        reexec_with_library_path()
        sys.path.insert(0, REF)
        import numpy as np
        # numpy >= 1.24 removed aliases MXNet 1.4's Python frontend still uses (np.bool etc.); restore them in THIS process only
        for name, typ in (("bool", bool), ("int", int), ("float", float), ("object", object), ("str", str)):
            if not hasattr(np, name):
                setattr(np, name, typ)
        import mxnet as mx
        from mxnet import autograd, gluon, nd
    except Exception as e:  # pragma: no cover - depends on the box
        print(json.dumps({"impl": "reference", "unavailable": "reference build failed to load: %r" % (e,)}))
        return 0

    N, B, K, W = args.gpus, args.batch_size, args.steps, max(3, args.warmup)
    ctxs = [mx.cpu(i) for i in range(N)] if args.cpu else [mx.gpu(i) for i in range(N)]
    kvtype = args.kvstore or ("local" if args.cpu else ("nccl" if N > 1 else "device"))

    # ---- the reference's model, verbatim from examples/cnn.py:56-64
    net = gluon.nn.Sequential()
    net.add(gluon.nn.Conv2D(channels=16, kernel_size=5, activation="relu"), gluon.nn.MaxPool2D(pool_size=2, strides=2),
            gluon.nn.Conv2D(channels=32, kernel_size=5, activation="relu"), gluon.nn.MaxPool2D(pool_size=2, strides=2),
            gluon.nn.Dense(256, activation="relu"), gluon.nn.Dense(128, activation="relu"), gluon.nn.Dense(10))
    net.initialize(mx.init.Xavier(), ctx=ctxs)
    loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()

    rng = np.random.RandomState(100)
    pool = 64
    Xh = [[nd.array(rng.rand(B, 1, 28, 28).astype("float32"), ctx=mx.cpu_pinned() if not args.cpu else mx.cpu()) for _ in ctxs] for _ in range(pool)]
    yh = [[nd.array(rng.randint(0, 10, (B,)).astype("float32"), ctx=mx.cpu_pinned() if not args.cpu else mx.cpu()) for _ in ctxs] for _ in range(pool)]
    net(Xh[0][0].as_in_context(ctxs[0]))          # materialise deferred shapes
    params = list(net.collect_params().values())
    assert sum(int(np.prod(p.shape)) for p in params) == 178762

    kv = mx.kv.create(kvtype)
    kv.set_optimizer(mx.optimizer.Adam(learning_rate=0.01))
    for idx, p in enumerate(params):
        kv.init(idx, p.data(ctxs[0]))
        kv.pull(idx, out=p.list_data(), priority=-idx)

    def step(xs, ys):
        with autograd.record():
            losses = [loss_fn(net(x), y) for x, y in zip(xs, ys)]
        for l in losses:
            l.backward()
        for idx, p in enumerate(params):          # examples/cnn.py:121-125
            kv.push(idx, [g / B for g in p.list_grad()], priority=-idx)
            kv.pull(idx, out=p.list_data(), priority=-idx)
        return losses

    dev_batches = [([x.as_in_context(c) for x, c in zip(Xh[i], ctxs)], [y.as_in_context(c) for y, c in zip(yh[i], ctxs)]) for i in range(2)]
    for i in range(W):
        step(*dev_batches[i % 2])
    nd.waitall()
    t0 = time.perf_counter()
    for i in range(K):
        step(*dev_batches[i % 2])
    nd.waitall()
    dev_s = time.perf_counter() - t0
    # end to end: H2D of every batch from (pinned) host memory + D2H of the loss every step
    nd.waitall()
    t0 = time.perf_counter()
    last = 0.0
    for i in range(K):
        j = (W + i) % pool
        xs = [x.as_in_context(c) for x, c in zip(Xh[j], ctxs)]
        ys = [y.as_in_context(c) for y, c in zip(yh[j], ctxs)]
        losses = step(xs, ys)
        last = float(sum(l.mean().asscalar() for l in losses) / len(losses))
    nd.waitall()
    e2e_s = time.perf_counter() - t0
    out = {
        "metric": "cnn.py samples/sec (whole box, device-timed, max over ranks)",
        "value": round(N * B * K / dev_s, 1), "unit": "samples/s", "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": round(dev_s / K * 1e3, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "impl": "reference",
        "config": {"model": "examples/cnn.py MNIST CNN (Conv16k5-Pool-Conv32k5-Pool-Dense256-Dense128-Dense10, 178762 params)",
                   "global_batch": B * N, "per_gpu_batch": B, "seq_len": None, "kvstore": kvtype,
                   "parallelism": "MXNet 1.4 single-process data parallel over %d device(s), per-key push/pull, Adam on the kvstore" % N,
                   "build": "unmodified /root/reference, make USE_CUDA=1 USE_CUDNN=0 USE_NCCL=1 USE_DIST_KVSTORE=0 USE_OPENCV=0 USE_LAPACK=0 "
                            "CUDA_ARCH=sm_100 (baseline/build_reference.sh)",
                   "timing": "host clock between mx.nd.waitall() barriers (MXNet exposes no CUDA events)", "mxnet": mx.__version__},
        "e2e": {"value": round(N * B * K / e2e_s, 1), "unit": "samples/s", "ms_per_step": round(e2e_s / K * 1e3, 5),
                "h2d_bytes_per_step": N * (B * 784 * 4 + B * 4), "d2h_bytes_per_step": N * 4, "final_loss": round(last, 5)},
        "gpu_launches": None,
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    try:
        rc = main()
    except SystemExit:
        raise
    except BaseException as e:  # pragma: no cover - depends on the box (e.g. a kvstore type the build cannot serve at this N)
        import traceback
        traceback.print_exc()
        print(json.dumps({"impl": "reference", "unavailable": "reference run failed: %r" % (e,)}))
        rc = 0
    sys.exit(rc)
