"""Phase timeline of the fused HiPS kernel on every rank (torchrun): stamps of CTA 0 — 0 start | 1 flags published | 2 party tiles done | 3 global tiles done | 4 first ready flag seen |
5 gradients cleared | 6 party flags seen | 7 tile loaded/reduced | 8 tile staged/applied | 9 arrivals seen | 10 global tile applied (last-iteration values) — for a few graph-replayed steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402
from geomx_b200.parallel import Topology  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
parties = int(sys.argv[1]) if len(sys.argv) > 1 else (2 if world % 2 == 0 else 1)
eng = mx.models.HipsCNNTrainStep(batch_size=32, topo=Topology(world, rank, parties, 1), device=dev, use_graph=True,
                                 use_multicast="--no-multicast" not in sys.argv)
X = torch.rand(32, 1, 28, 28).pin_memory(); y = torch.randint(0, 10, (32,)).float().pin_memory()
for _ in range(10):
    eng.step(X, y)
eng.fabric.state["fsa"][3] = 1
flush = torch.empty(64 * 1024 * 1024, device=dev)
for it in range(4):
    flush.fill_(1.0); eng.fabric.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.run_device(); b.record(); torch.cuda.synchronize()
    st = eng.fabric.state["fsa"][8:8 + 22].view(torch.int64).cpu().tolist()
    rel = ["%d:%.1f" % (i, (v - st[0]) / 1e3) for i, v in enumerate(st[:11]) if v]
    print("rank %d step %d: total %.1f us | hips phases (us since kernel start): %s" % (rank, it, a.elapsed_time(b) * 1e3, rel), flush=True)
    dist.barrier()
# isolated exchange latency: the fused kernel alone, back to back (gradients untouched), ranks aligned by the flag barrier
eng.fabric.state["fsa"][3] = 0
for reps in (1, 20):
    eng.fabric.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.fabric.fsa_step(zero_grad=True)
    b.record(); torch.cuda.synchronize()
    print("rank %d isolated hips x%d: %.1f us per launch" % (rank, reps, a.elapsed_time(b) * 1e3 / reps), flush=True)
    dist.barrier()
dist.destroy_process_group()
