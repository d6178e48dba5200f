"""GPU bring-up diagnostics: runs each tcgen05 GEMM configuration in its own subprocess (a hung mbarrier must not take the rest down)
and prints error statistics.  Usage: python tools/gpu_diag.py [case ...]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    "kk_small": (False, False, 128, 32, 32),
    "kk": (False, False, 256, 128, 512),
    "kk_odd": (False, False, 200, 72, 100),
    "k_mn": (False, True, 128, 64, 64),
    "mn_k": (True, False, 128, 64, 64),
    "mn_mn": (True, True, 128, 64, 64),
    "mn_mn_big": (True, True, 256, 512, 32),
    "k_mn_conv": (False, True, 2048, 400, 32),
}


def run_case(name):
    import torch
    from geomx_b200.ops import native
    torch.backends.cuda.matmul.allow_tf32 = False
    a_mn, b_mn, M, N, K = CASES[name]
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    D = torch.full((M, N), float("nan"), device="cuda")
    native.gemm(Am, Bm, D, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    ref = A @ B.t()
    err = (D - ref).abs()
    rel = float((D - ref).norm() / ref.norm())
    print("%-12s a_mn=%d b_mn=%d %dx%dx%d rel=%.3e max_abs=%.3e nan=%d" % (name, a_mn, b_mn, M, N, K, rel, float(err.nan_to_num(1e9).max()), int(D.isnan().sum())))
    if not rel < 2e-3:
        # localise: which 8x8 blocks are wrong
        bad = (err > 0.05 * ref.abs().max()).float()
        rows = bad.sum(1).nonzero().flatten()[:16].tolist(); cols = bad.sum(0).nonzero().flatten()[:16].tolist()
        print("   bad rows(first16)=%s bad cols(first16)=%s frac_bad=%.3f" % (rows, cols, float(bad.mean())))
        print("   D[0,:8]=%s\n   R[0,:8]=%s" % (D[0, :8].tolist(), ref[0, :8].tolist()))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2]); sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=90, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout.strip()[-1500:] if r.returncode else r.stdout.strip())
        except subprocess.TimeoutExpired:
            print("%-12s TIMEOUT (hung kernel)" % n)
