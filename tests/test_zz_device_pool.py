"""Native device-memory pool (csrc/kernels/storage_gpu.cu, ``mx.storage.DevicePool``).  (File name sorts last on purpose: the pluggable-allocator test swaps a process-wide allocator in a child process.)  The bucketing / reuse / reserve policy runs on a simulated
device (host malloc behind a capacity) so it is checked without a GPU; the CUDA backend and the PyTorch pluggable-allocator mode are GPU tests."""
import subprocess
import sys
import os

import pytest
import torch

import geomx_b200 as mx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MiB = 1 << 20


def test_naive_pool_buckets_reuse_and_streams():
    pool = mx.storage.DevicePool(-1, sim_capacity=64 * MiB, pool_type="Naive", page=4096, reserve=0)
    assert pool.round_size(1) == 4096 and pool.round_size(4097) == 8192 and pool.round_size(3 * MiB + 1) == 3 * MiB + 4096
    a = pool.alloc(5000, stream=1); b = pool.alloc(5000, stream=1)
    assert a != b and pool.stats() == {"used_bytes": 16384, "cached_bytes": 0, "num_alloc": 2, "num_pool_hits": 0, "num_driver_alloc": 2}
    assert pool.free(a, stream=1) is True
    assert pool.alloc(8000, stream=2) != a                       # same bucket, another stream: not handed out (no cross-stream reuse without a sync)
    c = pool.alloc(6000, stream=1)
    assert c == a and pool.stats()["num_pool_hits"] == 1         # same bucket, same stream: the cached block
    with pytest.raises(ValueError):
        pool.free(12345)
    pool.free(b, 1); pool.free(c, 1)
    assert pool.stats()["cached_bytes"] == 16384 and pool.stats()["used_bytes"] == 8192
    pool.release_all()
    assert pool.stats()["cached_bytes"] == 0
    d = pool.alloc(5000, stream=1)
    assert pool.stats()["num_driver_alloc"] == 4                 # the cache was really given back
    pool.free(d, 1)


def test_round_pool_and_large_buckets():
    pool = mx.storage.DevicePool(-2, sim_capacity=1 << 40, pool_type="Round", page=4096, reserve=0, cutoff=20)
    assert pool.round_size(100) == 4096 and pool.round_size(MiB) == MiB                     # linear below the cutoff
    assert pool.round_size(MiB + 1) == 2 * MiB and pool.round_size(5 * MiB) == 8 * MiB      # powers of two above it
    big = 65 << 30
    assert pool.round_size(big) == 72 << 30                                                 # above 1 GiB: sixteenths of the enclosing power of two, not 128 GiB
    a = pool.alloc(3 * MiB); pool.free(a)
    assert pool.alloc(4 * MiB) == a                                                         # both live in the 4 MiB bucket


def test_reserve_releases_the_cache_before_growing():
    pool = mx.storage.DevicePool(-3, sim_capacity=100 * MiB, pool_type="Naive", reserve=10)
    blocks = [pool.alloc(20 * MiB) for _ in range(4)]             # 80 MiB live: 20 MiB free, the reserve is 10 MiB
    for b in blocks:
        pool.free(b)
    assert pool.stats()["cached_bytes"] == 80 * MiB
    x = pool.alloc(30 * MiB)                                      # does not fit next to the cache + reserve: the cache goes first
    st = pool.stats()
    assert st["cached_bytes"] == 0 and st["used_bytes"] == 30 * MiB
    with pytest.raises(MemoryError):
        pool.alloc(90 * MiB)
    pool.free(x)
    y = pool.alloc(95 * MiB)                                      # an allocation that fails is retried after a release of everything cached
    assert pool.stats()["used_bytes"] == 95 * MiB
    pool.free(y)


def test_unpooled_frees_immediately():
    pool = mx.storage.DevicePool(-4, sim_capacity=10 * MiB, pool_type="Unpooled")
    a = pool.alloc(1000)
    assert pool.round_size(1000) == 1024 and pool.free(a) is False and pool.stats()["cached_bytes"] == 0
    with pytest.raises(ValueError):
        mx.storage.DevicePool(-5, sim_capacity=1, cutoff=99)


@pytest.mark.gpu
def test_cuda_backend_tensors_from_the_pool():
    pool = mx.storage.DevicePool(0)
    before = pool.stats()
    t = pool.empty(1000, torch.float32)
    assert t.is_cuda and t.numel() == 1000 and t.dtype == torch.float32
    t.fill_(3.0); u = t[10:20] * 2
    assert float(u.sum()) == 60.0
    ptr = t.data_ptr()
    del t
    torch.cuda.synchronize()
    st = pool.stats()
    assert st["cached_bytes"] - before["cached_bytes"] == pool.round_size(4000)
    t2 = pool.empty(900, torch.float32)                           # same bucket, same stream: the same block
    assert t2.data_ptr() == ptr and pool.stats()["num_pool_hits"] == before["num_pool_hits"] + 1
    del t2
    pool.release_all()
    assert pool.stats()["cached_bytes"] == 0


@pytest.mark.gpu
def test_native_pool_as_torch_allocator():
    """GEOMX_GPU_MEM_POOL=native: every CUDA tensor of the process comes from the native pool (a fresh process: the allocator can only be
    swapped before the first CUDA allocation)."""
    code = (
        "import torch, geomx_b200 as mx\n"
        "pool = mx.storage.DevicePool(0)\n"
        "x = torch.ones(1 << 20, device='cuda'); y = (x * 2).sum().item(); assert y == float(2 << 20)\n"
        "st = pool.stats(); assert st['num_alloc'] > 0 and st['used_bytes'] >= 4 << 20, st\n"
        "a = mx.nd.ones((64, 64), ctx=mx.gpu(0)); b = mx.nd.dot(a, a); assert float(b.asnumpy()[0, 0]) == 64.0\n"
        "w = torch.randn(256, 256, device='cuda', requires_grad=True); (w @ w).sum().backward(); assert w.grad is not None\n"
        "del x, a, b, w; torch.cuda.synchronize()\n"
        "st = pool.stats(); assert st['cached_bytes'] > 0 and st['num_pool_hits'] >= 0, st\n"
        "z = torch.empty(1 << 20, device='cuda'); assert pool.stats()['num_pool_hits'] >= 1\n"
        "print('NATIVE_POOL_OK')\n")
    env = dict(os.environ, GEOMX_GPU_MEM_POOL="native", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NATIVE_POOL_OK" in r.stdout, r.stdout + r.stderr
