"""Tiny driver for `ncu --set full`: runs the dense0-shaped tcgen05 GEMM (M=32,N=256,K=512, bias+ReLU), the conv1-shaped GEMM and one
fused HiPS step so that a single short process contains the top kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomx_b200.ops import native  # noqa: E402

torch.manual_seed(0)
x = torch.randn(32, 512, device="cuda"); w = torch.randn(256, 512, device="cuda"); b = torch.randn(256, device="cuda")
y = torch.empty(32, 256, device="cuda")
col = torch.randn(2048, 400, device="cuda"); wc = torch.randn(32, 400, device="cuda"); z = torch.empty(32, 32, 8, 8, device="cuda")
A = torch.randn(8192, 4096, device="cuda"); Bm = torch.randn(4096, 4096, device="cuda"); D = torch.empty(8192, 4096, device="cuda")
for _ in range(3):
    native.gemm(x, w, y, bias=b, relu=True)
    native.gemm(col, wc, z, bias=b[:32], relu=True, store_nchw_hw=64)
    native.gemm(A, Bm, D)
torch.cuda.synchronize()
print("ok", float(y.sum()), float(D.sum()))
