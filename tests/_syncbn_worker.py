"""Worker for the SyncBatchNorm tests: each rank normalises its share of a batch with SyncBatchNorm; rank 0 compares with BatchNorm on the
whole.  Backend: gloo on CPU (default) or NCCL with one GPU per rank (SYNCBN_DEVICE=cuda: the statistics cross NVLink)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402

on_gpu = os.environ.get("SYNCBN_DEVICE", "cpu") == "cuda"
if on_gpu:
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
else:
    dist.init_process_group("gloo")
ctx = mx.gpu(int(os.environ.get("LOCAL_RANK", 0))) if on_gpu else mx.cpu()
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.RandomState(0)
X = rng.randn(8, 3, 4, 4).astype(np.float32) * 2 + 1
W = rng.randn(8, 3, 4, 4).astype(np.float32)
part = slice(rank * 8 // world, (rank + 1) * 8 // world)


def run(layer, xs, ws):
    layer.initialize(ctx=ctx)
    x = mx.nd.array(xs, ctx=ctx); x.attach_grad()
    with mx.autograd.record():
        y = layer(x)
        loss = (y * mx.nd.array(ws, ctx=ctx)).sum()
    loss.backward()
    return y.asnumpy(), x.grad.asnumpy(), layer.gamma.grad().asnumpy(), layer.running_var.data().asnumpy()


y, dx, dg, rv = run(mx.gluon.contrib.nn.SyncBatchNorm(in_channels=3), X[part], W[part])
dgt = torch.from_numpy(dg.copy())
if on_gpu:
    dgt = dgt.cuda()
dist.all_reduce(dgt)
dgt = dgt.cpu()
if rank == 0:
    y0, dx0, dg0, rv0 = run(mx.gluon.nn.BatchNorm(in_channels=3), X, W)
    tol = 2e-4 if on_gpu else 1e-5        # the single-device reference on a GPU is the native BatchNorm kernel (different summation order)
    ok = np.allclose(y, y0[part], atol=tol) and np.allclose(dx, dx0[part], atol=tol) and np.allclose(dgt.numpy(), dg0, atol=10 * tol, rtol=1e-4) \
        and np.allclose(rv, rv0, atol=tol, rtol=1e-4)
    print("SYNCBN", "PASS" if ok else "FAIL", float(np.abs(y - y0[part]).max()), float(np.abs(dx - dx0[part]).max()),
          float(np.abs(dgt.numpy() - dg0).max()), float(np.abs(rv - rv0).max()))
dist.barrier()
dist.destroy_process_group()
