"""``mx.engine`` — the host-side dependency engine (versioned variables, read/write ordering, priorities).

Parity: ``include/mxnet/engine.h:95-314`` / ``src/engine/threaded_engine*.{h,cc}`` / ``naive_engine.cc`` and ``python/mxnet/engine.py``
(``bulk``).  Device work is ordered by CUDA streams in this design; the engine (native: ``csrc/runtime/engine.h``) schedules HOST work with
the reference's semantics — ops that only read a variable run concurrently, a writer waits for earlier readers/writers, ready ops are served
by priority, by per-device worker pools (compute / copy per GPU, normal / priority on the CPU); exceptions surface at the wait points.  Selected by ``MXNET_ENGINE_TYPE`` (``NaiveEngine`` = run inline, for debugging races; default threaded with
``MXNET_CPU_WORKER_NTHREADS`` workers).  Consumers: asynchronous checkpoint writes (``mx.nd.save_async``), user callbacks;
``mx.nd.waitall()`` drains it."""
from __future__ import annotations

import contextlib
import os

from . import runtime

__all__ = ["get", "push", "new_variable", "delete_variable", "wait_for_var", "wait_all", "stats", "bulk", "set_bulk_size", "engine_type",
           "NORMAL", "COPY", "PRIORITY"]

NORMAL, COPY, PRIORITY = 0, 1, 2      # FnProperty (include/mxnet/engine.h:59-77): which pool of the op's device runs it

_engine = None
_bulk = 0


def engine_type():
    return os.environ.get("MXNET_ENGINE_TYPE", "ThreadedEnginePerDevice")


def get():
    """The process-wide engine (created on first use)."""
    global _engine
    if _engine is None:
        if not runtime.available():
            raise RuntimeError("native runtime not built")
        naive = engine_type() == "NaiveEngine"
        _engine = runtime.C().Engine(int(os.environ.get("MXNET_CPU_WORKER_NTHREADS", "2")), naive)
    return _engine


def new_variable():
    return get().new_variable()


def _device_of(ctx):
    if ctx is None:
        return -1
    if isinstance(ctx, int):
        return ctx
    return ctx.device_id if getattr(ctx, "device_type", "cpu") == "gpu" else -1


def push(fn, const_vars=(), mutable_vars=(), priority=0, name="", ctx=None, prop=NORMAL):
    """Schedule ``fn()`` once every earlier writer of ``const_vars`` and every earlier reader/writer of ``mutable_vars`` has finished.

    ``ctx`` (a Context or a GPU index; default CPU) and ``prop`` choose the worker pool: every device has its own compute pool and its own
    copy pool (``COPY``), the CPU has a normal and a ``PRIORITY`` pool — host work for one GPU never queues behind host work for another
    (``ThreadedEnginePerDevice``).  An exception raised by ``fn`` is remembered on ``mutable_vars`` and re-raised by the next
    ``wait_for_var`` on one of them, or by ``wait_all``."""
    get().push(fn, list(const_vars), list(mutable_vars), int(priority), name, _device_of(ctx), int(prop))


def delete_variable(var):
    """Forget ``var`` once every op pushed so far that touches it has finished."""
    get().delete_variable(var)


def stats():
    """Ops executed so far per worker pool: ``{"cpu": n, "priority": n, "gpu0": n, "gpu0/copy": n, ...}``."""
    return dict(get().stats()) if _engine is not None else {}


def wait_for_var(var):
    get().wait_for_var(var)


def wait_all():
    if _engine is not None:
        _engine.wait_for_all()


def set_bulk_size(size):
    """Kept for API compatibility (bulk execution concerns device-op fusion, which CUDA graphs cover here); returns the previous value."""
    global _bulk
    prev, _bulk = _bulk, int(size)
    return prev


@contextlib.contextmanager
def bulk(size):
    prev = set_bulk_size(size)
    try:
        yield
    finally:
        set_bulk_size(prev)
