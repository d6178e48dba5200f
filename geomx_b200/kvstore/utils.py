"""Row-id dedup for row_sparse pulls.  Parity: ``src/kvstore/kvstore_utils.{cc,cu}`` ``UniqueImpl`` (CUB radix sort
+ ``DeviceSelect::Unique``; blocking D2H of the count).  CUDA ids go through the native ``gx_unique_i64`` (csrc/kernels/sparse_ops.cu,
the same CUB pair), host ids through ``torch.unique``."""
import torch


def unique_rows(ids: torch.Tensor) -> torch.Tensor:
    ids = ids.reshape(-1).long()
    if ids.is_cuda:
        from ..ops import native
        if native.available():
            from ..ops import _native_api
            return _native_api.unique_i64(ids)
    return torch.unique(ids, sorted=True)
