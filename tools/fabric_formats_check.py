"""Wire formats of the fused HiPS step on real GPUs (torchrun, >= 2 ranks): fp16 transport and Bi-Sparse between the tiers, checked
against a plain PyTorch fp32 oracle assembled from all-gathered gradients.  Server optimizer: none (the aggregate is what workers pull —
the cnn_bsc.py / cnn_fp16.py flow with a local Trainer)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402,F401
from geomx_b200.parallel import Topology  # noqa: E402
from geomx_b200.parallel.arena import ArenaLayout  # noqa: E402
from geomx_b200.parallel.fabric import HipsFabric  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    parties = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    topo = Topology(world, rank, parties, 1)
    S, P = topo.party_size, topo.num_parties
    layout = ArenaLayout.build([(0, (300,)), (1, (64, 100)), (2, (5000,)), (3, (7,)), (4, (40, 128)), (5, (3000,))])
    f = HipsFabric(layout, topo, dev, None)
    f.set_push_scale(0.5)
    ok = True
    thr = 0.02
    K = int(1024 * thr)
    f.set_wire_formats({1: "bsc", 2: "fp16", 4: "bsc", 5: "fp8"}, thr)
    dgt = "--dgt" in sys.argv
    if dgt:
        f.enable_dgt(k=0.5, alpha=0.3)
    ema = torch.zeros(f.tiles, device=dev)
    n = f.n
    u = torch.zeros(P, n, device=dev); v = torch.zeros(P, n, device=dev)       # oracle copies of every party's residual state
    tile_fmt = f.tile_fmt.cpu().numpy()
    def q8(x):                      # block-scaled e4m3: one scale per 128 values (a warp's share of a tile)
        b = x.view(-1, 128)
        scale = b.abs().amax(1, keepdim=True).clamp_min(1e-30) * (1.0 / 448.0)
        return ((b / scale).to(torch.float8_e4m3fn).float() * scale).view(-1)

    for step in range(5 if dgt else 3):
        tile_fmt = f.tile_fmt.cpu().numpy()
        g = torch.Generator(device="cpu").manual_seed(1000 * step + rank)
        grad = torch.randn(n, generator=g).to(dev)
        if step == 2:
            grad = grad * (torch.rand(n, generator=g).to(dev) < 0.05)          # very sparse gradients: residuals dominate
        f.grad.tensor.copy_(grad)
        allg = [torch.empty_like(grad) for _ in range(world)]
        dist.all_gather(allg, grad)
        f.fsa_step(zero_grad=True)
        torch.cuda.synchronize()
        # ---------------- oracle
        expect = torch.zeros(n, device=dev)
        tol = torch.zeros(n, device=dev)
        for t in range(f.tiles):
            sl = slice(t * 1024, (t + 1) * 1024)
            fmt = int(tile_fmt[t])
            agg_parties = []
            for gp in range(P):
                owner_local = t % S
                acc = allg[gp * S + owner_local][sl].clone()
                for j in range(S):
                    if j != owner_local:
                        x = allg[gp * S + j][sl]
                        acc += x.half().float() if fmt == 1 else (q8(x) if fmt == 3 else x)
                agg_parties.append(acc * 0.5)
            if fmt == 0:
                expect[sl] = sum(agg_parties)
                gagg = expect[sl]
            elif fmt == 3:
                w = sum(q8(a) for a in agg_parties)
                expect[sl] = w.half().float()
                gagg = w
                tol[sl] = 0.07 * w.abs().max() + 1e-6                            # at most one e4m3 rounding step of difference per hop
            elif fmt == 1:
                w = sum(a.half().float() for a in agg_parties)
                expect[sl] = w.half().float()
                gagg = w
                tol[sl] = 2e-3 * w.abs().max() + 1e-6                            # summation-order / double-rounding slack
            else:
                out = torch.zeros(1024, device=dev)
                for gp in range(P):
                    uu = 0.9 * u[gp, sl] + agg_parties[gp]
                    vv = v[gp, sl] + uu
                    idx = torch.topk(vv.abs(), K).indices
                    out[idx] += vv[idx]
                    uu[idx] = 0; vv[idx] = 0
                    u[gp, sl] = uu; v[gp, sl] = vv
                expect[sl] = out
                tol[sl] = 1e-6
                gagg = out
            m = gagg.abs().mean()
            ema[t] = m if float(ema[t]) == 0.0 else 0.3 * ema[t] + 0.7 * m
        got = f.param.tensor
        bad = ((got - expect).abs() > tol + 1e-6 * expect.abs()).sum().item()
        nz = int((got[torch.from_numpy(np.repeat(tile_fmt, 1024) == 2).to(dev)] != 0).sum())
        zeroed = bool((f.grad.tensor == 0).all())
        print("rank %d step %d: mismatches=%d  bsc non-zeros pulled=%d (<= %d)  grads cleared=%s  protocol_err=%s" % (
            rank, step, bad, nz, int((tile_fmt == 2).sum()) * P * K, zeroed, f.check_protocol_errors()), flush=True)
        ok = ok and bad == 0 and zeroed and not f.check_protocol_errors()
        if dgt and step % 2 == 1:
            c = f.dgt_rerank()
            cerr = float(((c - ema).abs() / (ema.abs() + 1e-12)).max())
            order = f.tile_order.long()
            sorted_ok = bool((c[order][:-1] >= c[order][1:]).all()) and sorted(order.tolist()) == list(range(f.tiles))
            fmt_now = f.tile_fmt.cpu().numpy()
            n_imp = max(1, int(round(0.5 * f.tiles)))
            demoted = [int(fmt_now[t]) for t in order[n_imp:].tolist()]
            print("rank %d step %d: dgt contrib max rel err %.2e  order sorted=%s  demoted fmts=%s" % (rank, step, cerr, sorted_ok, sorted(set(demoted))), flush=True)
            ok = ok and cerr < 1e-3 and sorted_ok and all(x in (2, 3) for x in demoted)
        dist.barrier()
    t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("FORMATS_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
