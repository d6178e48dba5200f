"""In-graph cost of every kernel of the flagship step: captures CUDA graphs of the first k launches (k = 1..15) and reports the marginal
time of each (device-timed, L2 flushed between replays like bench.py).  Usage: python tools/step_breakdown.py [--no-flush]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402
from geomx_b200.parallel import Topology  # noqa: E402

flush_on = "--no-flush" not in sys.argv
from geomx_b200.ops import native  # noqa: E402
native.set_gemm_precision("tf32" if "--fast" in sys.argv else "3xtf32")
eng = mx.models.HipsCNNTrainStep(batch_size=32, topo=Topology(1, 0, 1, 1), use_graph=False)
eng.x.copy_(torch.rand(32, 1, 28, 28)); eng.label.copy_(torch.randint(0, 10, (32,)).float())
for _ in range(3):
    eng._body()
torch.cuda.synchronize()
names = [s[0] for s in eng._steps()]
flush = torch.empty(64 * 1024 * 1024, device="cuda")
prev = 0.0
print("# precision=%s fused_mlp=%s overlap=%s flush=%s" % (native.gemm_precision(), eng.fused_mlp, eng.overlap, flush_on))
print("%-48s %10s %10s" % ("kernel", "cum us", "marginal"))
for k in range(1, len(names) + 1):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng._body(stop_after=k)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        eng._body(stop_after=k)
    ts = []
    for _ in range(60):
        if flush_on:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print("%-48s %10.2f %10.2f" % (names[k - 1], med, med - prev))
    prev = med
