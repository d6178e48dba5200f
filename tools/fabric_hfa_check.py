"""HFA on the fabric (torchrun, >= 2 GPUs): party rounds (LL kernel in party mode or the flag kernel) and global rounds through the per-key
KVStore API, against the closed form: after a party round every member holds the party mean of what was pushed, after a global round the
mean of the party means."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MXNET_KVSTORE_USE_HFA"] = "1"
os.environ.setdefault("MXNET_KVSTORE_HFA_K2", "2")
import geomx_b200 as mx  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if len(sys.argv) > 1:
        os.environ["GEOMX_NUM_PARTIES"] = sys.argv[1]
    kv = mx.kv.create("dist_sync")
    topo = kv._topo
    shapes = [(300,), (40, 50), (7,)]
    params = [mx.nd.zeros(s, ctx=mx.gpu(int(os.environ.get("LOCAL_RANK", 0)))) for s in shapes]
    for i, p in enumerate(params):
        kv.init(i, p)
    ok = True
    S, P = topo.party_size, topo.num_parties
    for it in range(1, 5):
        vals = []
        for i, p in enumerate(params):
            local = mx.nd.array(np.full(shapes[i], float(it * 10 + rank + i), dtype=np.float32), ctx=p.context)
            kv.push(i, local / kv.num_workers)
            kv.pull(i, p)
        mx.nd.waitall()
        for i, p in enumerate(params):
            party_mean = lambda g: np.mean([it * 10 + r + i for r in range(g * S, (g + 1) * S)])
            expect = party_mean(topo.party) if it % 2 else np.mean([party_mean(g) for g in range(P)])
            got = float(p.asnumpy().reshape(-1)[0]); same = bool((p.asnumpy() == got).all())
            if abs(got - expect) > 1e-4 or not same:
                ok = False
                print("rank %d iter %d key %d: got %.5f expected %.5f" % (rank, it, i, got, expect), flush=True)
    t = torch.tensor([1 if ok and not kv.fabric.check_protocol_errors() else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("protocol", kv.fabric.protocol, "parties", P, "party_size", S)
        print("HFA_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
