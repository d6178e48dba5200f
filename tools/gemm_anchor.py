"""Large-GEMM anchor: gx_gemm (tcgen05, TF32 and 3xTF32) next to cuBLAS on the same box, same clocks, same shape.

  python tools/gemm_anchor.py [M N K]            (default 8192 4096 4096)

Every variant: 5 warm-up calls, then 20 calls between two CUDA events (inputs 3 x 128 MiB > L2 is not needed here: the working set of one
call, 320 MiB, already exceeds the 126 MB L2).  The SM clock is sampled with NVML during each timed loop.  Prints TFLOP/s and the relative
error against an fp64 product of a 512-row slice."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomx_b200.ops import native  # noqa: E402


class Clocks:
    def __init__(self):
        self.vals, self.stop = [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h, self.nv = pynvml.nvmlDeviceGetHandleByIndex(0), pynvml
        except Exception:
            self.nv = None

    def __enter__(self):
        if self.nv:
            self.t = threading.Thread(target=self._run); self.t.start()
        return self

    def _run(self):
        while not self.stop:
            self.vals.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            time.sleep(0.002)

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join()

    def median(self):
        v = sorted(self.vals)
        return v[len(v) // 2] if v else None


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks() as c:
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, c.median()


def main():
    M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8192, 4096, 4096)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); D = torch.empty(M, N, device=dev)
    Ab, Bb = A.bfloat16(), B.bfloat16()
    ref = A[:512].double() @ B.double().t()
    rel = lambda X: float((X[:512].double() - ref).norm() / ref.norm())
    flop = 2.0 * M * N * K
    rows = []

    def add(name, fn, out):
        ms, mhz = timed(fn)
        rows.append((name, ms, flop / ms / 1e9, rel(out()), mhz))

    native.set_gemm_precision("tf32")
    add("gx_gemm tcgen05 TF32", lambda: native.gemm(A, B, D), lambda: D)
    native.set_gemm_precision("3xtf32")
    add("gx_gemm tcgen05 3xTF32 (fp32-accurate)", lambda: native.gemm(A, B, D), lambda: D)
    torch.backends.cuda.matmul.allow_tf32 = True
    Bt = B.t()
    add("cuBLAS TF32 (torch.matmul, allow_tf32)", lambda: torch.matmul(A, Bt, out=D), lambda: D)
    torch.backends.cuda.matmul.allow_tf32 = False
    add("cuBLAS fp32 (torch.matmul)", lambda: torch.matmul(A, Bt, out=D), lambda: D)
    Db = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    Bbt = Bb.t()
    add("cuBLAS bf16 (torch.matmul)", lambda: torch.matmul(Ab, Bbt, out=Db), lambda: Db.float())
    print("GEMM %d x %d x %d  (A [M,K] row-major, B [N,K] row-major, D = A.B^T)" % (M, N, K))
    print("%-42s %9s %10s %12s %8s" % ("variant", "ms", "TFLOP/s", "rel err fp64", "SM MHz"))
    for name, ms, tf, e, mhz in rows:
        print("%-42s %9.3f %10.1f %12.2e %8s" % (name, ms, tf, e, mhz))
    base = rows[2][2]
    print("gx TF32 / cuBLAS TF32 = %.2f   gx 3xTF32 / cuBLAS fp32 = %.2f" % (rows[0][2] / base, rows[1][2] / rows[3][2]))


if __name__ == "__main__":
    main()
