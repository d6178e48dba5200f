"""Device contexts (``mx.cpu()``, ``mx.gpu(i)``).

Parity: ``python/mxnet/context.py`` in the reference.  A Context maps 1:1 onto a
``torch.device``; ``gpu(i)`` raises :class:`MXNetError` lazily (on first
allocation) when no CUDA device exists, which is what ``examples/utils.py:try_gpu``
relies on.
"""
from __future__ import annotations

import threading

import torch

from .base import MXNetError

__all__ = ["Context", "cpu", "gpu", "cpu_pinned", "cpu_shared", "current_context", "num_gpus"]


class Context:
    devtype2str = {1: "cpu", 2: "gpu", 3: "cpu_pinned", 5: "cpu_shared"}
    devstr2type = {"cpu": 1, "gpu": 2, "cuda": 2, "cpu_pinned": 3, "cpu_shared": 5}
    _default = threading.local()

    def __init__(self, device_type, device_id=0):
        if isinstance(device_type, Context):
            self.device_typeid, self.device_id = device_type.device_typeid, device_type.device_id
        else:
            self.device_typeid = Context.devstr2type[str(device_type)]
            self.device_id = int(device_id)
        self._old = None

    @property
    def device_type(self) -> str:
        return Context.devtype2str[self.device_typeid]

    @property
    def torch_device(self) -> torch.device:
        if self.device_typeid == 2:
            return torch.device("cuda", self.device_id)
        return torch.device("cpu")

    def place(self, t: torch.Tensor) -> torch.Tensor:
        """Give ``t`` (already on ``torch_device``) the storage kind of this context: ``cpu_pinned`` = page-locked host memory (asynchronous
        H2D / D2H copies; reference ``Context::CPUPinned``, src/storage/pinned_memory_storage.h), ``cpu_shared`` = POSIX shared memory that
        other processes can map (reference ``Context::CPUShared``, src/storage/cpu_shared_storage_manager.h — the DataLoader hand-off)."""
        if self.device_typeid == 3 and t.device.type == "cpu" and torch.cuda.is_available() and not t.is_pinned():
            return t.pin_memory()
        if self.device_typeid == 5 and t.device.type == "cpu" and not t.is_shared():
            return t.share_memory_()
        return t

    def check_available(self):
        if self.device_typeid == 2:
            if not torch.cuda.is_available() or self.device_id >= torch.cuda.device_count():
                raise MXNetError("gpu(%d) is not available (no CUDA device)" % self.device_id)

    def __hash__(self):
        return hash((self.device_typeid, self.device_id))

    def __eq__(self, other):
        return isinstance(other, Context) and self.device_typeid == other.device_typeid \
            and self.device_id == other.device_id

    def __str__(self):
        return "%s(%d)" % (self.device_type, self.device_id)

    __repr__ = __str__

    def __enter__(self):
        self._old = getattr(Context._default, "value", None)
        Context._default.value = self
        return self

    def __exit__(self, *a):
        Context._default.value = self._old

    @property
    def default_ctx(self):
        """The context entered with ``with ctx:`` in this thread, if any (python/mxnet/context.py: Context._default_ctx.value)."""
        return getattr(Context._default, "value", None) or Context("cpu", 0)

    @staticmethod
    def from_torch(dev: torch.device) -> "Context":
        if dev.type == "cuda":
            return Context("gpu", dev.index or 0)
        return Context("cpu", 0)


def cpu(device_id=0):
    return Context("cpu", device_id)


def cpu_pinned(device_id=0):
    return Context("cpu_pinned", device_id)


def cpu_shared(device_id=0):
    return Context("cpu_shared", device_id)


def gpu(device_id=0):
    return Context("gpu", device_id)


def num_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def current_context():
    v = getattr(Context._default, "value", None)
    return v if v is not None else Context("cpu", 0)


def gpu_memory_info(device_id=0):
    """``(free, total)`` bytes of a GPU (context.py:260-285)."""
    import torch
    if not torch.cuda.is_available():
        from .base import MXNetError
        raise MXNetError("gpu_memory_info: no GPU is visible to this process")
    return tuple(int(v) for v in torch.cuda.mem_get_info(device_id))
