"""``mx.util`` — small helpers (parity: python/mxnet/util.py: ``makedirs``, ``get_gpu_count``, ``get_gpu_memory``)."""
from __future__ import annotations

import os


def makedirs(d):
    os.makedirs(os.path.expanduser(d), exist_ok=True)


def get_gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def get_gpu_memory(gpu_dev_id):
    """``(free, total)`` bytes of a GPU."""
    import torch
    free, total = torch.cuda.mem_get_info(gpu_dev_id)
    return free, total
