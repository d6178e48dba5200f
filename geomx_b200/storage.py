"""Host storage pool + resources (``mx.storage``): pooled (optionally pinned) host buffers exposed as torch tensors.

Parity: ``src/storage/storage.cc`` / ``pooled_storage_manager.h`` (size-bucketed pools, ``MXNET_*_MEM_POOL_*`` knobs, ``ReleaseAll``) and
``src/resource.cc`` (temp workspaces, per-device RNG seeds).  The pool itself is native (``csrc/runtime/storage.h``); blocks are page-locked
through ``cudaHostRegister`` when a CUDA device is present so that the kvstore's staging buffers copy asynchronously.

Device memory: ``DevicePool`` is the native pooled manager (``csrc/kernels/storage_gpu.cu``: page / power-of-two buckets per (device, stream),
``MXNET_GPU_MEM_POOL_TYPE`` = Naive | Round | Unpooled, ``MXNET_GPU_MEM_POOL_RESERVE``, ``..._PAGE_SIZE``, ``..._ROUND_LINEAR_CUTOFF``).  By default
tensors come from PyTorch's caching allocator; ``use_native_gpu_pool()`` (or ``GEOMX_GPU_MEM_POOL=native`` at import) installs the native pool
as PyTorch's allocator for the process, ``DevicePool(dev).empty(...)`` hands out single tensors from it."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import runtime
from .base import getenv_int

__all__ = ["HostPool", "DevicePool", "use_native_gpu_pool", "default_pool", "pinned_empty", "seed", "next_seed"]

_DSIZE = {torch.float32: 4, torch.float64: 8, torch.float16: 2, torch.bfloat16: 2, torch.uint8: 1, torch.int8: 1, torch.int32: 4, torch.int64: 8}


_NP = {torch.float32: np.float32, torch.float64: np.float64, torch.float16: np.float16, torch.uint8: np.uint8, torch.int8: np.int8,
       torch.int32: np.int32, torch.int64: np.int64}


class HostPool:
    def __init__(self, page=4096, max_pooled=None, pin=None):
        if not runtime.available():
            raise RuntimeError("native runtime not built")
        C = runtime.C()
        self._pool = C.PooledHostStorage(page, max_pooled if max_pooled is not None else getenv_int("GEOMX_HOST_POOL_MAX_MB", 4096) << 20)
        self._res = C.ResourceManager()
        self._pin = torch.cuda.is_available() if pin is None else pin
        self._registered = set()

    def _register(self, ptr, size):
        if not self._pin or ptr in self._registered:
            return
        try:
            rt = torch.cuda.cudart()
            if int(rt.cudaHostRegister(ptr, size, 0)) == 0:
                self._registered.add(ptr)
        except Exception:
            pass

    def _unregister(self, ptr):
        if ptr in self._registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(ptr)
            except Exception:
                pass
            self._registered.discard(ptr)

    def empty(self, numel, dtype=torch.float32):
        """A 1-D tensor of ``numel`` elements backed by a pool block; the block returns to the pool when the tensor is collected."""
        nbytes = max(1, int(numel) * _DSIZE[dtype])
        ptr, _hit = self._pool.alloc(nbytes)
        if not ptr:
            raise MemoryError("host pool: cannot allocate %d bytes" % nbytes)
        size = self._pool.size_of(ptr)
        self._register(ptr, size)
        pool = self

        class _Owner:                                   # returns the block when the LAST tensor/view over it is gone
            def __del__(self_inner):
                if not pool._pool.free(ptr):
                    pool._unregister(ptr)
        raw = (ctypes.c_char * size).from_address(ptr)
        raw._owner = _Owner()                           # numpy keeps `raw` as the array base, torch keeps the array: lifetime = storage lifetime
        if numel == 0:
            return torch.empty(0, dtype=dtype)
        npdt = _NP.get(dtype)
        if npdt is None:                                # bfloat16: no numpy dtype, reinterpret 16-bit words
            return torch.from_numpy(np.frombuffer(raw, dtype=np.int16, count=int(numel))).view(dtype)
        return torch.from_numpy(np.frombuffer(raw, dtype=npdt, count=int(numel)))

    def is_pinned(self, t):
        return t.data_ptr() in self._registered or any(t.data_ptr() >= p and t.data_ptr() < p + self._pool.size_of(p) for p in self._registered)

    def release_all(self):
        for ptr, _ in self._pool.release_all():
            self._unregister(ptr)

    def stats(self):
        return dict(self._pool.stats())

    # resources
    def seed(self, s): self._res.seed(int(s))
    def next_seed(self, device=0): return int(self._res.next_seed(int(device)))


_default = None


def default_pool():
    global _default
    if _default is None:
        _default = HostPool()
    return _default


def pinned_empty(numel, dtype=torch.float32):
    return default_pool().empty(numel, dtype)


def seed(s): default_pool().seed(s)
def next_seed(device=0): return default_pool().next_seed(device)


# ------------------------------------------------------------------------------------------------------------ device memory
class DevicePool:
    """Native pooled device memory of one GPU (reference ``GPUPooledStorageManager`` / ``GPUPooledRoundedStorageManager``).

    ``device >= 0``: the CUDA device (policy from the ``MXNET_GPU_MEM_POOL_*`` variables, one pool per device and process).  ``device < 0`` with
    ``sim_capacity``: a simulated device backed by host malloc that runs the same bucketing / reserve / release policy without a GPU."""
    _TYPES = {"Naive": 0, "Round": 1, "Unpooled": 2}

    def __init__(self, device=0, sim_capacity=None, pool_type="Naive", page=4096, reserve=5, cutoff=24):
        from .ops import native
        self._lib = native.require()
        self.device = int(device)
        if self.device < 0:
            if sim_capacity is None:
                raise ValueError("a simulated device (device < 0) needs sim_capacity")
            if self._lib.gx_gpu_pool_create_sim(self.device, int(sim_capacity), self._TYPES[pool_type], int(page), int(reserve), int(cutoff)) != 0:
                raise ValueError("bad simulated pool parameters")

    def alloc(self, nbytes, stream=0):
        p = self._lib.gx_gpu_pool_alloc(self.device, int(nbytes), ctypes.c_void_p(int(stream)))
        if not p:
            raise MemoryError("device pool %d: cannot allocate %d bytes" % (self.device, nbytes))
        return p

    def free(self, ptr, stream=0):
        rc = self._lib.gx_gpu_pool_free(self.device, ctypes.c_void_p(ptr), ctypes.c_void_p(int(stream)))
        if rc < 0:
            raise ValueError("pointer %#x does not belong to device pool %d" % (ptr, self.device))
        return rc == 0                      # True: cached for reuse, False: released to the driver

    def round_size(self, nbytes):
        return int(self._lib.gx_gpu_pool_round_size(self.device, int(nbytes)))

    def release_all(self):
        self._lib.gx_gpu_pool_release_all(self.device)

    def stats(self):
        out = (ctypes.c_uint64 * 5)()
        self._lib.gx_gpu_pool_stats(self.device, out)
        return dict(zip(("used_bytes", "cached_bytes", "num_alloc", "num_pool_hits", "num_driver_alloc"), (int(v) for v in out)))

    def empty(self, numel, dtype=torch.float32):
        """A 1-D device tensor over a pool block; the block returns to the pool when the last view of the tensor dies."""
        if self.device < 0:
            raise RuntimeError("simulated devices hand out addresses, not tensors")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nbytes = max(1, int(numel)) * _DSIZE[dtype]
        ptr = self.alloc(nbytes, stream)
        pool = self

        class _Block:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}

            def __del__(self_inner):
                pool.free(ptr, stream)
        return torch.as_tensor(_Block(), device="cuda:%d" % self.device).view(dtype)[:int(numel)]


_native_allocator = None


def use_native_gpu_pool():
    """Install the native pool as PyTorch's CUDA allocator (``torch.cuda.memory.CUDAPluggableAllocator``).  Must run before the first CUDA
    allocation of the process; blocks are cached per (device, stream), so a tensor handed to another stream needs the usual synchronisation."""
    global _native_allocator
    if _native_allocator is None:
        from .ops import native
        native.require()
        _native_allocator = torch.cuda.memory.CUDAPluggableAllocator(native._LIB_PATH, "gx_torch_alloc", "gx_torch_free")
        torch.cuda.memory.change_current_allocator(_native_allocator)
    return _native_allocator
