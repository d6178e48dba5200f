"""``mx.engine`` — the host-side dependency engine (versioned variables, read/write ordering, priorities).

Parity: ``include/mxnet/engine.h:95-314`` / ``src/engine/threaded_engine*.{h,cc}`` / ``naive_engine.cc`` and ``python/mxnet/engine.py``
(``bulk``).  Device work is ordered by CUDA streams in this design; the engine (native: ``csrc/runtime/engine.h``) schedules HOST work with
the reference's semantics — ops that only read a variable run concurrently, a writer waits for earlier readers/writers, ready ops are served
by priority.  Selected by ``MXNET_ENGINE_TYPE`` (``NaiveEngine`` = run inline, for debugging races; default threaded with
``MXNET_CPU_WORKER_NTHREADS`` workers).  Consumers: asynchronous checkpoint writes (``mx.nd.save_async``), user callbacks;
``mx.nd.waitall()`` drains it."""
from __future__ import annotations

import contextlib
import os

from . import runtime

__all__ = ["get", "push", "new_variable", "wait_for_var", "wait_all", "bulk", "set_bulk_size", "engine_type"]

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


def push(fn, const_vars=(), mutable_vars=(), priority=0, name=""):
    """Schedule ``fn()`` once every earlier writer of ``const_vars`` and every earlier reader/writer of ``mutable_vars`` has finished."""
    get().push(fn, list(const_vars), list(mutable_vars), int(priority), name)


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
