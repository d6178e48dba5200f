"""Row-id dedup for row_sparse pulls.  Parity: ``src/kvstore/kvstore_utils.{cc,cu}`` ``UniqueImpl`` (CUB radix sort
+ ``DeviceSelect::Unique``; blocking D2H of the count).  Here: ``torch.unique`` (CCCL radix sort + unique underneath on
CUDA) — kept as a library call because row_sparse is off every GeoMX config path."""
import torch


def unique_rows(ids: torch.Tensor) -> torch.Tensor:
    return torch.unique(ids.reshape(-1).long(), sorted=True)
