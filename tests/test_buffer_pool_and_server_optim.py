"""Message-buffer pool (csrc/hips/block_pool.h) and the multi-threaded native server optimizer (csrc/hips/server_optim.h)."""
import numpy as np
import pytest

from geomx_b200 import runtime

_C = runtime.C()


def test_size_classes_are_monotonic_with_bounded_slack():
    assert _C.buffer_pool_class_of(1) == 64 << 10 and _C.buffer_pool_class_of(64 << 10) == 64 << 10
    assert _C.buffer_pool_class_of((64 << 10) + 1) == 128 << 10 and _C.buffer_pool_class_of(1 << 20) == 1 << 20
    prev = 0
    for n in [70_000, 300_000, (1 << 20) + 1, 3 << 20, (5 << 20) + 17, 64 << 20, (1 << 30) + 5]:
        c = _C.buffer_pool_class_of(n)
        assert c >= n and c >= prev and (c - n) <= max(n, 64 << 10) * (1.0 if n <= (1 << 20) else 0.26), (n, c)
        prev = c


def test_released_blocks_are_reused_and_limit_is_respected():
    _C.buffer_pool_trim()
    limit = _C.buffer_pool_stats()["limit_bytes"]
    try:
        _C.buffer_pool_set_limit(64 << 20)
        before = _C.buffer_pool_stats()
        seen = _C.buffer_pool_probe(3 << 20, 12)          # allocate-and-drop, like a stream of received frames
        st = _C.buffer_pool_stats()
        assert len(seen) == 1 and st["hits"] - before["hits"] == 11 and st["misses"] - before["misses"] == 1
        assert st["cached_bytes"] == _C.buffer_pool_class_of(3 << 20)
        # small arrays bypass the pool
        _C.buffer_pool_probe(1000, 5)
        assert _C.buffer_pool_stats()["hits"] == st["hits"] and _C.buffer_pool_stats()["misses"] == st["misses"]
        # a block larger than the limit is never cached
        _C.buffer_pool_probe(80 << 20, 2)
        assert _C.buffer_pool_stats()["cached_bytes"] == st["cached_bytes"]
        # limit 0: pooling off, cache trimmed
        _C.buffer_pool_set_limit(0)
        assert _C.buffer_pool_stats()["cached_bytes"] == 0
        h = _C.buffer_pool_stats()["hits"]
        _C.buffer_pool_probe(3 << 20, 4)
        assert _C.buffer_pool_stats()["hits"] == h and _C.buffer_pool_stats()["cached_bytes"] == 0
    finally:
        _C.buffer_pool_set_limit(limit)
        _C.buffer_pool_trim()


def _ref(spec, w, grads):
    """numpy fp32 reference of MXNet's sgd(_mom)_update / adam_update / DCASGD, step by step."""
    f = np.float32
    w = w.astype(f).copy(); a = np.zeros_like(w); b = np.zeros_like(w)
    kind = spec["name"]
    lr0, wd, rs, clip = f(spec.get("lr", 0.01)), f(spec.get("wd", 0)), f(spec.get("rescale_grad", 1)), f(spec.get("clip_gradient", -1))
    mom, b1, b2, eps, lam = f(spec.get("momentum", 0)), f(spec.get("beta1", 0.9)), f(spec.get("beta2", 0.999)), f(spec.get("epsilon", 1e-8)), f(spec.get("lamda", 0.04))
    if kind == "dcasgd":
        b = w.copy()
    for t, g in enumerate(grads, 1):
        g = g.astype(f) * rs
        if kind == "adam":
            g = g + wd * w
        if clip >= 0:
            g = np.clip(g, -clip, clip)
        if kind == "adam":
            lr = f(lr0 * np.sqrt(f(1) - f(b2) ** f(t)) / (f(1) - f(b1) ** f(t)))
            a = b1 * a + (f(1) - b1) * g
            b = b2 * b + (f(1) - b2) * g * g
            w = w - lr * a / (np.sqrt(b) + eps)
        elif kind == "sgd":
            g = g + wd * w
            if mom != 0:
                a = mom * a - lr0 * g; w = w + a
            else:
                w = w - lr0 * g
        else:
            upd = g + wd * w + lam * g * g * (w - b)
            prev = w.copy()
            if mom != 0:
                a = mom * a - lr0 * upd; w = w + a
            else:
                w = w - lr0 * upd
            b = prev
    return w


@pytest.mark.parametrize("spec", [
    {"name": "sgd", "lr": 0.05, "wd": 1e-4},
    {"name": "sgd", "lr": 0.05, "momentum": 0.9, "clip_gradient": 0.5, "rescale_grad": 0.5},
    {"name": "adam", "lr": 0.01, "wd": 1e-3, "clip_gradient": 1.0},
    {"name": "dcasgd", "lr": 0.02, "momentum": 0.8, "lamda": 0.04},
])
@pytest.mark.parametrize("n", [1000, (1 << 20) + 37])          # below the threading grain / split over several threads with a ragged tail
def test_native_optimizer_matches_numpy_on_small_and_threaded_sizes(spec, n):
    rng = np.random.default_rng(n % 1000 + len(spec))
    w = rng.standard_normal(n).astype(np.float32)
    grads = rng.standard_normal((3, n)).astype(np.float32)
    s = ";".join("%s=%s" % kv for kv in spec.items())
    got = np.asarray(_C.native_optimizer_run(s, w, grads))
    want = _ref(spec, w, grads)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    assert _C.server_threads() >= 1
