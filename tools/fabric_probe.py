"""Latency of the fabric primitives on this box (torchrun, >= 2 GPUs): local / peer / multimem loads, stores + fence.sys, flag hand-off.
Also prints `nvidia-smi topo -m` so the numbers can be read against the link type.  Output goes into profiles/ via tools/make_profiles.py."""
import ctypes
import os
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomx_b200 as mx  # noqa: E402,F401
from geomx_b200.ops import native  # noqa: E402
from geomx_b200.parallel import Topology  # noqa: E402
from geomx_b200.parallel.fabric import SymmetricHeap  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
heap = SymmetricHeap(Topology(world, rank, 1, 1), dev)
TILES = 256
buf = heap.alloc(TILES * 1024, torch.float32)
flg = heap.alloc(8192, torch.int32)
peer = (rank + 1) % world
out = torch.zeros(64, dtype=torch.int64, device=dev)
lib = native.require()
NAMES = ["local weak ld.v4", "local ld.relaxed.sys", "peer weak ld.v4", "peer ld.relaxed.sys", "multimem.ld_reduce", "peer st + fence.sys",
         "multimem.st + fence.sys", "fence.sys (idle)", "peer st, bar, 1 fence, flag st", "local ld.acquire.sys", "local st + threadfence"]
if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout)
    except Exception as e:
        print("nvidia-smi topo failed", e)
    print("backend", heap.backend, "multicast_ptr", hex(buf.multicast_ptr))
for grid in (1, 132):
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        rc = lib.gx_fabric_probe(buf.tensor.data_ptr(), buf.peer_ptrs[peer], buf.multicast_ptr or None, flg.peer_ptrs[peer], out.data_ptr(), TILES, grid,
                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
        torch.cuda.synchronize()
    o = out.cpu().tolist()
    if rank == 0:
        print("grid=%d CTAs x 256 threads x 16 B (ns, 1st / 2nd access):" % grid)
        for k, n in enumerate(NAMES):
            print("  %-34s %7d %7d" % (n, o[2 * k], o[2 * k + 1]))
dist.barrier()
dist.destroy_process_group()
