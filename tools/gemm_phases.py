"""Where does a small tcgen05 GEMM spend its time?  Chains NCHAIN dense0-shaped GEMMs in a CUDA graph (PDL edges) and prints (a) the
per-GEMM time for several TMA ring depths, (b) %globaltimer stamps of the phases of CTA (0,0,0):
0 kernel start | 1 prologue done | 2 griddepcontrol.wait returned | 3 first TMA stage landed | 4 last MMA committed | 5 accumulator visible | 6 epilogue done"""
import ctypes
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomx_b200.ops import native  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "--sweep":
    for st in ("2", "4", "8"):
        for pdl in ("1", "0"):
            env = dict(os.environ, GEOMX_GEMM_STAGES=st, GEOMX_PDL=pdl)
            out = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
            print("stages=%s pdl=%s : %s" % (st, pdl, out.strip().splitlines()[-1] if out.strip() else "?"))
    sys.exit(0)

NCHAIN = 12
x = torch.randn(32, 512, device="cuda"); w = torch.randn(512, 512, device="cuda") * 0.05; b = torch.zeros(512, device="cuda")
bufs = [torch.empty(32, 512, device="cuda") for _ in range(2)]
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
lib = native.require()


def chain():
    src = x
    for i in range(NCHAIN):
        dst = bufs[i & 1]
        native.gemm(src, w, dst, bias=b, relu=True)
        src = dst


for _ in range(3):
    chain()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain()
ts = []
for _ in range(50):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); e.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(e) * 1e3)
ts.sort()
# phase stamps of the LAST gemm of an eager chain (debug pointer is baked at launch time)
lib.gx_gemm_set_debug(ctypes.c_void_p(dbg.data_ptr()))
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    chain()
lib.gx_gemm_set_debug(ctypes.c_void_p(0))
g2.replay(); torch.cuda.synchronize()
st = dbg.cpu().tolist()
rel = [(v - st[0]) / 1e3 for v in st[:7]]
print("chain of %d GEMMs (32x512x512): %.2f us total, %.2f us per GEMM | last-GEMM phases (us since its start): %s" % (
    NCHAIN, ts[len(ts) // 2], ts[len(ts) // 2] / NCHAIN, ["%.2f" % r for r in rel]))
