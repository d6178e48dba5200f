"""Multi-GPU correctness check of the fused HiPS kernels (run under torchrun, one rank per GPU).

  torchrun --nproc-per-node N tools/fabric_check.py [--parties P] [--gs G] [--no-multicast] [--mode dist_sync|dist_async]

Checks (all ranks): (1) with SGD(lr) on the global shard, w' == w - lr * (sum over ranks of grad)/B, where the per-rank gradients are
all-reduced with NCCL as the oracle; (2) every rank holds identical parameters after every step; (3) CUDA-graph replay of the step
trains (loss decreases) and keeps ranks consistent."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402
from geomx_b200.parallel import Topology  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parties", type=int, default=0)
    ap.add_argument("--gs", type=int, default=0, help="global servers; 0 = every rank, ownership sharded tile by tile (fabric default)")
    ap.add_argument("--no-multicast", action="store_true")
    ap.add_argument("--mode", default="dist_sync")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    parties = a.parties or (2 if world % 2 == 0 else 1)
    topo = Topology(world, rank, parties, a.gs)
    B = 32
    torch.manual_seed(7)
    eng = mx.models.HipsCNNTrainStep(batch_size=B, optimizer=mx.optimizer.SGD(learning_rate=0.1), topo=topo, device=dev, use_graph=False,
                                     use_multicast=not a.no_multicast, mode=a.mode, fused_zero_grad=False)
    if rank == 0:
        print("fabric backend=%s protocol=%s multicast=%s parties=%d party_size=%d gs=%s channels=%s" % (
            eng.fabric.heap.backend, eng.fabric.protocol, eng.fabric.use_multicast, parties, topo.party_size, topo.gs_ranks,
            {k: ("replicated" if v["replicate"] else "sharded") for k, v in eng.fabric.channels.items()}), flush=True)
    g = torch.Generator().manual_seed(100 + rank)
    X = torch.rand(B, 1, 28, 28, generator=g).to(dev); y = torch.randint(0, 10, (B,), generator=g).float().to(dev)
    eng.x.copy_(X); eng.label.copy_(y)
    ok = True
    for it in range(3):
        w0 = eng.fabric.param.tensor.clone()
        eng._body()
        torch.cuda.synchronize()
        gsum = eng.fabric.grad.tensor.clone()
        dist.all_reduce(gsum)
        expect = w0 - 0.1 * gsum / B
        err = float((eng.fabric.param.tensor - expect).abs().max())
        ref = eng.fabric.param.tensor.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, eng.fabric.param.tensor))
        if a.mode == "dist_async":
            # MixedSync: every party applies its own aggregate to the global owner's master weights on arrival and pulls what is there at that
            # moment.  SGD updates commute, so after the first round each TILE this rank pulled is either  w0 - lr*g_own_party/B  (the other
            # party had not arrived yet) or  w0 - lr*(g_own_party + g_other_parties)/B  — nothing else is a legal value.
            if it == 0:
                gparty = eng.fabric.grad.tensor.clone()
                # party sum through the world all-reduce: zero the other parties' contributions
                contrib = [torch.zeros_like(gparty) for _ in range(topo.num_parties)]
                contrib[topo.party].copy_(gparty)
                for c in contrib:
                    dist.all_reduce(c)
                own = w0 - 0.1 * contrib[topo.party] / B
                full = w0 - 0.1 * sum(contrib) / B
                p = eng.fabric.param.tensor
                dev_own, dev_full = (p - own).abs(), (p - full).abs()
                legal = torch.minimum(dev_own, dev_full)
                if topo.num_parties > 2:        # any subset of the other parties may have arrived: bound by the envelope instead
                    lo, hi = torch.minimum(own, full), torch.maximum(own, full)
                    legal = torch.clamp(lo - p, min=0) + torch.clamp(p - hi, min=0)
                worst = float(legal.max())
                print("rank %d iter 0 async: max distance to a legal value %.3e (own-party-only tiles: %d of %d)" % (
                    rank, worst, int((dev_own < dev_full).sum()), p.numel()), flush=True)
                ok = ok and worst < 1e-5
            else:
                print("rank %d iter %d async max|w-sync expectation|=%.3e" % (rank, it, err), flush=True)
        else:
            print("rank %d iter %d max|w-expect|=%.3e identical_to_rank0=%s loss=%.4f" % (rank, it, err, same, float(eng.loss.mean())), flush=True)
            ok = ok and err < 1e-5 and same
        dist.barrier()
    # graph mode + Adam
    eng2 = mx.models.HipsCNNTrainStep(batch_size=B, optimizer=mx.optimizer.Adam(learning_rate=0.01), topo=topo, device=dev, use_graph=True,
                                      use_multicast=not a.no_multicast, mode=a.mode)
    Xh, yh = X.cpu().pin_memory(), y.cpu().pin_memory()
    l0 = eng2.step(Xh, yh)
    for _ in range(60):
        l = eng2.step(Xh, yh)
    ref = eng2.fabric.param.tensor.clone(); dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, eng2.fabric.param.tensor))
    print("rank %d graph: loss %.4f -> %.4f identical=%s" % (rank, l0, l, same), flush=True)
    lm = torch.tensor([l0, l], device=dev); dist.all_reduce(lm)   # every rank trains on its own random batch: judge the job-wide loss
    ok = ok and float(lm[1]) < float(lm[0]) and (same or a.mode == "dist_async") and not eng.fabric.check_protocol_errors() and not eng2.fabric.check_protocol_errors()
    t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("FABRIC_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
