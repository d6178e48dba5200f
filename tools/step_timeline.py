"""In-graph timeline of the flagship step on every rank (run under torchrun, or plain python for 1 GPU).

Every kernel of the step writes %globaltimer stamps (debug buffers of cnn_direct.cu / mlp_chain.cu, phase stamps of the exchange channels); the
graph is captured with stamping on and replayed like bench.py does (L2 flushed, ranks re-aligned before each step).  Prints, per rank, the
median start / end of each kernel relative to the start of the first one — where the step's microseconds go at N GPUs.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_timeline.py [--lookahead]
"""
import ctypes
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402
from geomx_b200.ops import native  # noqa: E402
from geomx_b200.parallel import Topology  # noqa: E402


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    parties = 2 if world >= 2 and world % 2 == 0 else 1
    topo = Topology(world, rank, parties, 0)
    native.set_gemm_precision("3xtf32")
    lib = native.require()
    cdbg = torch.zeros(1024, dtype=torch.int64, device=dev)
    mdbg = torch.zeros(32, dtype=torch.int64, device=dev)
    lib.gx_cnn_set_debug(ctypes.c_void_p(cdbg.data_ptr()))
    lib.gx_mlp_chain_set_debug(ctypes.c_void_p(mdbg.data_ptr()))
    torch.manual_seed(1)
    eng = mx.models.HipsCNNTrainStep(batch_size=32, optimizer=mx.optimizer.Adam(learning_rate=0.01), topo=topo, device=dev, use_graph=True,
                                     lookahead="--lookahead" in sys.argv)
    X = torch.rand(32, 1, 28, 28).pin_memory(); y = torch.randint(0, 10, (32,)).float().pin_memory()
    for _ in range(5):
        eng.step(X, y)
    fab = eng.fabric
    chans = list(fab.channels) or ["fsa"]
    for c in chans:
        fab.state[c][3] = 1
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    rows = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(24):
        flush.fill_(float(it))
        if world > 1:
            fab.barrier()
        ev[0].record(); eng.run_device(); ev[1].record()
        torch.cuda.synchronize()
        c, m = cdbg.tolist(), mdbg.tolist()
        st = {k: fab.state[k][8:8 + 12].view(torch.int64).tolist() for k in chans}
        row = {"cnn_fwd": (c[0], c[4]), "mlp_chain": (m[0], m[9]), "cnn_bwd": (c[8], c[13]), "cnn_wgrad1": (c[16], c[18])}
        for k in chans:
            if st[k][0] > 0:
                row["chan:" + k] = (st[k][0], st[k][5])
        if getattr(eng, "fused_exchange", False):       # exchange tail of the backward launch: 20 ticket | 21 grid done | 22 sent | 23 received | 24 end
            row["bwd tail"] = (c[20], c[24])
            row["_tail"] = [(v - c[20]) / 1e3 for v in c[20:25]]
        if not getattr(eng, "fused_exchange", False):
            per = [(c[64 + 3 * i], c[65 + 3 * i], c[66 + 3 * i]) for i in range(256)]
            row["_cta"] = per
        row["_ms"] = ev[0].elapsed_time(ev[1])
        row["_ph"] = {k: [(v - st[k][0]) / 1e3 for v in st[k][:6]] for k in chans if st[k][0] > 0}
        rows.append(row)
    rows = rows[4:]
    names = [k for k in rows[0] if not k.startswith("_")]
    lines = ["rank %d  step %.2f us (event-timed median)" % (rank, 1e3 * statistics.median(r["_ms"] for r in rows))]
    for k in sorted(names, key=lambda k: statistics.median(r[k][0] - min(v[0] for kk, v in r.items() if not kk.startswith("_") and v[0] > 0) for r in rows)):
        rel0 = [(r[k][0] - min(v[0] for kk, v in r.items() if not kk.startswith("_") and v[0] > 0)) / 1e3 for r in rows]
        rel1 = [(r[k][1] - min(v[0] for kk, v in r.items() if not kk.startswith("_") and v[0] > 0)) / 1e3 for r in rows]
        lines.append("  %-12s start %7.2f  end %7.2f  (dur %6.2f us)" % (k, statistics.median(rel0), statistics.median(rel1),
                                                                        statistics.median(b - a for a, b in zip(rel0, rel1))))
    if "_cta" in rows[0]:
        # how the 256 CTAs of the convolution-backward grid pack onto the SMs (last sampled step)
        r = rows[-1]
        base = min(v[0] for kk, v in r.items() if not kk.startswith("_") and v[0] > 0)
        for name, lo, hi in (("bwd CTAs  (0..127)", 0, 128), ("wgrad CTAs (128..255)", 128, 256)):
            st = sorted((a - base) / 1e3 for a, b, sm_ in r["_cta"][lo:hi]); en = sorted((b - base) / 1e3 for a, b, sm_ in r["_cta"][lo:hi])
            du = sorted((b - a) / 1e3 for a, b, sm_ in r["_cta"][lo:hi])
            lines.append("  %-22s start min/med/max %.2f %.2f %.2f | end %.2f %.2f %.2f | duration %.2f %.2f %.2f" % (
                name, st[0], st[64], st[-1], en[0], en[64], en[-1], du[0], du[64], du[-1]))
        sms = {}
        for i, (a, b, sm_) in enumerate(r["_cta"]):
            sms.setdefault(sm_, []).append(i)
        shared = sorted(len(v) for v in sms.values())
        lines.append("  SMs used %d; CTAs per SM: %s" % (len(sms), {k: shared.count(k) for k in sorted(set(shared))}))
        pair = {"bwd+bwd": 0, "bwd+wgrad": 0, "wgrad+wgrad": 0}
        for v in sms.values():
            if len(v) == 2:
                k = sum(1 for i in v if i >= 128)
                pair[["bwd+bwd", "bwd+wgrad", "wgrad+wgrad"][k]] += 1
        lines.append("  SM sharing: %s" % pair)
    if "_tail" in rows[0]:
        lines.append("  bwd tail (us since first tail ticket): ticket|grid done|sent|received|end %s" % ["%.2f" % statistics.median(r["_tail"][i] for r in rows) for i in range(5)])
    for k in rows[0]["_ph"]:
        lines.append("  chan:%s phases since its start (us): %s" % (k, ["%.2f" % statistics.median(r["_ph"][k][i] for r in rows) for i in range(6)]))
    for r in range(world):
        if r == rank:
            print("\n".join(lines), flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
