"""KVStore API surface on the fabric (torchrun, >= 2 GPUs): Python-executed optimizers, priorities, row_sparse_pull, 2-bit compression.

Every check compares the pulled values with a closed form computed from the known per-rank gradients."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "hostopt"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ctx = mx.gpu(int(os.environ.get("LOCAL_RANK", 0)))
    kv = mx.kv.create("dist_sync")
    ok = True
    shapes = [(300,), (40, 50), (7,)]
    if mode == "hostopt":
        # AdaGrad has no native spec: w -= lr * g / (sqrt(h) + eps), h += g^2, executed by the Python updater on the aggregated gradient
        kv.set_optimizer(mx.optimizer.AdaGrad(learning_rate=0.5, eps=1e-7))
    elif mode == "sched":
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=1.0, lr_scheduler=mx.lr_scheduler.FactorScheduler(step=1, factor=0.5, base_lr=1.0)))
    elif mode == "2bit":
        kv.set_gradient_compression({"type": "2bit", "threshold": 0.5})
    params = [mx.nd.ones(s, ctx=ctx) for s in shapes]
    emb = mx.nd.array(np.arange(40, dtype=np.float32).reshape(10, 4), ctx=ctx)
    for i, p in enumerate(params):
        kv.init(i, p)
        kv.pull(i, p)
    if mode == "rowsparse":
        kv.init(50, emb)            # the fabric lays its arena out at the first data operation: every key is initialised before that
        kv.pull(50, emb)
    mx.nd.waitall()
    gsum = sum(range(1, world + 1))                       # rank r pushes (r + 1) * base
    if mode in ("hostopt", "sched"):
        w = [np.ones(s, dtype=np.float64) for s in shapes]
        h = [np.zeros(s, dtype=np.float64) for s in shapes]
        for step in range(3):
            for i, p in enumerate(params):
                base = 0.1 * (i + 1)
                kv.push(i, mx.nd.array(np.full(shapes[i], base * (rank + 1), dtype=np.float32), ctx=ctx), priority=-i)
                kv.pull(i, p, priority=-i)
            mx.nd.waitall()
            for i in range(len(shapes)):
                g = 0.1 * (i + 1) * gsum
                if mode == "hostopt":
                    h[i] += g * g
                    w[i] -= 0.5 * g / (np.sqrt(h[i]) + 1e-7)
                else:
                    lr = 1.0 * 0.5 ** max(0, step)      # FactorScheduler(step=1): lr halves after every update (num_update = step + 1)
                    w[i] -= lr * g
                got = params[i].asnumpy().astype(np.float64)
                if np.abs(got - w[i]).max() > 1e-4:
                    ok = False
                    print("rank %d mode %s step %d key %d: got %.6f expected %.6f" % (rank, mode, step, i, got.reshape(-1)[0], w[i].reshape(-1)[0]), flush=True)
    elif mode == "2bit":
        # thr 0.5: a gradient of 0.3 sends 0 and keeps 0.3; the second push sends +0.5 (residual 0.6 -> 0.1); no optimizer: pull = aggregate
        outs = []
        for step in range(2):
            kv.push(0, mx.nd.array(np.full(shapes[0], 0.3, dtype=np.float32), ctx=ctx))
            kv.pull(0, params[0]); mx.nd.waitall()
            outs.append(float(params[0].asnumpy()[0]))
        if abs(outs[0] - 0.0) > 1e-6 or abs(outs[1] - 0.5 * world) > 1e-6:
            ok = False
            print("rank %d 2bit: got %s expected [0, %.2f]" % (rank, outs, 0.5 * world), flush=True)
    elif mode == "rowsparse":
        kv2 = kv
        out = mx.nd.sparse.zeros("row_sparse", (10, 4), ctx=ctx)
        kv2.row_sparse_pull(50, out=out, row_ids=mx.nd.array([7, 2, 7, 5], ctx=ctx, dtype="int64"))
        ids = out.indices.asnumpy().tolist(); rows = out.data.asnumpy()
        if ids != [2, 5, 7] or not np.allclose(rows[:, 0], [8.0, 20.0, 28.0]):
            ok = False
            print("rank %d row_sparse_pull: ids %s rows %s" % (rank, ids, rows[:, 0]), flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("API_CHECK", mode, "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
