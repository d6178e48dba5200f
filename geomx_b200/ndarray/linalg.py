"""``mx.nd.linalg`` — the linear-algebra operator namespace (reference: ``python/mxnet/ndarray/linalg.py``, populated from the ``_linalg_*``
operators of ``src/operator/tensor/la_op.cc``).  All operators work on the last two axes and broadcast over leading batch axes."""
import torch

from . import op_lib as _ops
from .ndarray import NDArray

__all__ = ["gemm", "gemm2", "potrf", "potri", "trmm", "trsm", "sumlogdiag", "syrk", "gelqf", "syevd", "extractdiag", "makediag", "extracttrian",
           "maketrian", "det", "slogdet", "inverse"]

gemm, gemm2, potrf, potri, trmm, trsm = _ops.linalg_gemm, _ops.linalg_gemm2, _ops.linalg_potrf, _ops.linalg_potri, _ops.linalg_trmm, _ops.linalg_trsm
sumlogdiag, syrk, gelqf, syevd = _ops.linalg_sumlogdiag, _ops.linalg_syrk, _ops.linalg_gelqf, _ops.linalg_syevd
extractdiag, makediag, det, inverse = _ops.linalg_extractdiag, _ops.linalg_makediag, _ops.linalg_det, _ops.linalg_inverse


def _tri_index(n, offset, lower, device):
    if offset > 0:
        lower = False
    elif offset < 0:
        lower = True
    return (torch.tril_indices if lower else torch.triu_indices)(n, n, offset, device=device)


def extracttrian(A, offset=0, lower=True):
    """The triangle of each matrix (row-major order of its entries) as a vector: ``(..., n, n) -> (..., n(n+1)/2)`` for ``offset=0``."""
    t = A._t
    r, c = _tri_index(t.shape[-1], offset, lower, t.device)
    return NDArray(t[..., r, c])


def maketrian(A, offset=0, lower=True):
    """Inverse of :func:`extracttrian`: scatter a vector of ``m(m+1)/2`` entries into the triangle of an ``(m+|offset|)``-square matrix."""
    t = A._t
    k = t.shape[-1]
    m = int((-1 + (1 + 8 * k) ** 0.5) / 2 + 0.5)
    n = m + abs(offset)
    r, c = _tri_index(n, offset, lower, t.device)
    out = torch.zeros(t.shape[:-1] + (n, n), dtype=t.dtype, device=t.device)
    out[..., r, c] = t
    return NDArray(out)


def slogdet(A):
    """``(sign, log|det|)`` of each matrix."""
    s, l = torch.linalg.slogdet(A._t)
    return NDArray(s), NDArray(l)
