"""Round-2 bring-up diagnostics (1 GPU): which step configuration survives CUDA-graph capture, GEMM accuracy vs fp64."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
import geomx_b200 as mx
from geomx_b200.parallel import Topology
torch.manual_seed(0)
eng = mx.models.HipsCNNTrainStep(batch_size=32, topo=Topology(1, 0, 1, 0), use_graph=True)
X = torch.rand(32, 1, 28, 28).pin_memory(); y = torch.randint(0, 10, (32,)).float().pin_memory()
l0 = eng.step(X, y)
for _ in range(40): l = eng.step(X, y)
print("RESULT loss %%.4f -> %%.4f launches/step %%d" %% (l0, l, eng.kernels_per_step))
''' % ROOT

def run(env):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    tail = [l for l in r.stdout.splitlines() if l.strip()][-1:] 
    print(env, "rc=%d" % r.returncode, tail[0][:200] if tail else "", flush=True)

if __name__ == "__main__":
    for env in ({"GEOMX_FUSED_MLP": "0", "GEOMX_STEP_OVERLAP": "0"}, {"GEOMX_FUSED_MLP": "1", "GEOMX_STEP_OVERLAP": "0"},
                {"GEOMX_FUSED_MLP": "0", "GEOMX_STEP_OVERLAP": "1"}, {"GEOMX_FUSED_MLP": "1", "GEOMX_STEP_OVERLAP": "1"},
                {"GEOMX_FUSED_MLP": "1", "GEOMX_STEP_OVERLAP": "1", "GEOMX_PDL": "0"},
                {"GEOMX_FUSED_MLP": "1", "GEOMX_STEP_OVERLAP": "0", "GEOMX_MLP_PDL": "0"}):
        run(env)
    import torch
    sys.path.insert(0, ROOT)
    from geomx_b200.ops import native as nat
    torch.manual_seed(0)
    for (M, N, K) in ((32, 256, 512), (2048, 32, 400), (128, 256, 32), (1024, 1024, 1024), (8192, 4096, 4096)):
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); D = torch.zeros(M, N, device="cuda")
        ref64 = (A.double() @ B.double().t())
        torch.backends.cuda.matmul.allow_tf32 = False
        ref32 = A @ B.t()
        rel = lambda x: float((x.double() - ref64).norm() / ref64.norm())
        out = {}
        for mode in ("3xtf32", "tf32"):
            nat.set_gemm_precision(mode); nat.gemm(A, B, D); torch.cuda.synchronize(); out[mode] = rel(D)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for _ in range(3): nat.gemm(A, B, D)
            ev[0].record()
            for _ in range(10): nat.gemm(A, B, D)
            ev[1].record(); torch.cuda.synchronize()
            out[mode + "_us"] = ev[0].elapsed_time(ev[1]) * 100
        nat.set_gemm_precision("3xtf32")
        print("GEMM %dx%dx%d rel-err vs fp64: 3xtf32 %.3e  tf32 %.3e  torch-fp32 %.3e | us/call 3xtf32 %.1f tf32 %.1f | TFLOP/s %.1f / %.1f" % (
            M, N, K, out["3xtf32"], out["tf32"], rel(ref32), out["3xtf32_us"], out["tf32_us"],
            2.0 * M * N * K / out["3xtf32_us"] / 1e6, 2.0 * M * N * K / out["tf32_us"] / 1e6), flush=True)
