"""``mx.profiler`` — chrome://tracing profiler with remote (server) control.

Parity: ``python/mxnet/profiler.py`` (set_config / set_state / dump / dumps / pause / resume, Domain / Task / Frame /
Event / Counter / Marker; ``profile_process='server'`` routes through the kvstore handle :28-33,63-66) and
``src/profiler/profiler.{h,cc}`` (chrome-trace JSON ``traceEvents`` with ``ph`` codes, pid = device index :155-254;
aggregate table ``aggregate_stats.cc``; continuous dump :258-296).

Design: two recorders with one file format.  Worker-side events (this module) are appended to an in-process list — host ranges use
a monotonic clock, device ranges use CUDA events that are resolved lazily at dump time, so recording never synchronises a stream;
hooks: every gluon ``Block`` call (``profile_imperative``), every symbolic graph node (``profile_symbolic``), kvstore push / pull,
plus user ``scope`` / ``Domain`` objects.  Server processes have no Python on their hot path and record with the native profiler
(``csrc/runtime/profiler.h``: push / pull handlers, controlled remotely through command 6, files prefixed ``rank<r>_``).
"""
from __future__ import annotations

import json
import os
import threading
import time

import torch

__all__ = ["set_config", "profiler_set_config", "set_state", "profiler_set_state", "dump", "dump_profile", "dumps",
           "pause", "resume", "Domain", "Task", "Frame", "Event", "Counter", "Marker", "scope", "set_kvstore_handle"]

_cfg = {"filename": "profile.json", "profile_all": False, "profile_symbolic": False, "profile_imperative": False,
        "profile_memory": False, "profile_api": False, "aggregate_stats": False, "continuous_dump": False,
        "dump_period": 1.0}
_state = {"running": False, "paused": False}
_events = []
_lock = threading.Lock()
_dev_ranges = []           # (name, cat, start_evt, end_evt, host_ts_us_at_start, device_index)
_kv_handle = None
_t0 = time.perf_counter()


def _now_us():
    return (time.perf_counter() - _t0) * 1e6


def set_kvstore_handle(handle):
    global _kv_handle
    _kv_handle = handle


def set_config(**kwargs):
    """Accepts the reference's keys; ``profile_process='server'`` forwards to all servers via the kvstore.  ``continuous_dump=True`` rewrites
    the trace file every ``dump_period`` seconds while the profiler runs (profiler.cc:258-296)."""
    proc = kwargs.pop("profile_process", "worker")
    if proc == "server":
        assert _kv_handle is not None, "create a dist kvstore before configuring the server profiler"
        _kv_handle.set_server_profiler_command(0, ",".join("%s:%s" % (k, v) for k, v in kwargs.items()))
        return
    _cfg.update(kwargs)
    if _cfg["continuous_dump"]:
        _start_continuous_dump()


_dump_thread = None


def _start_continuous_dump():
    global _dump_thread
    if _dump_thread is not None and _dump_thread.is_alive():
        return

    def loop():
        while _cfg["continuous_dump"]:
            time.sleep(max(0.05, float(_cfg["dump_period"])))
            if _state["running"]:
                dump(finished=False)
    _dump_thread = threading.Thread(target=loop, name="profiler-continuous-dump", daemon=True)
    _dump_thread.start()


profiler_set_config = set_config


def set_state(state="stop", profile_process="worker"):
    if profile_process == "server":
        assert _kv_handle is not None
        _kv_handle.set_server_profiler_command(1, str(int(state == "run")))
        return
    _state["running"] = state == "run"


profiler_set_state = set_state


def pause(profile_process="worker"):
    if profile_process == "server":
        _kv_handle.set_server_profiler_command(2, "1"); return
    _state["paused"] = True


def resume(profile_process="worker"):
    if profile_process == "server":
        _kv_handle.set_server_profiler_command(2, "0"); return
    _state["paused"] = False


def is_active():
    return _state["running"] and not _state["paused"]


def _emit(ev):
    with _lock:
        _events.append(ev)


def _resolve_device_ranges():
    out = []
    for name, cat, s, e, host_us, dev in _dev_ranges:
        try:
            e.synchronize()
            dur = s.elapsed_time(e) * 1e3
        except RuntimeError:
            continue
        out.append({"name": name, "cat": cat, "ph": "X", "ts": host_us, "dur": dur, "pid": dev, "tid": "stream"})
    _dev_ranges.clear()
    return out


def dumps(reset=False, format="table"):
    with _lock:
        evs = list(_events) + _resolve_device_ranges()
        if reset:
            _events.clear()
    agg, open_ = {}, {}

    def add(name, dur):
        a = agg.setdefault(name, [0, 0.0, float("inf"), 0.0])
        a[0] += 1; a[1] += dur; a[2] = min(a[2], dur); a[3] = max(a[3], dur)
    for e in evs:
        ph = e.get("ph")
        if ph == "X":
            add(e["name"], e["dur"])
        elif ph == "B":                                      # Task / Frame / Event durations are begin/end pairs
            open_.setdefault((e["name"], e.get("cat")), []).append(e["ts"])
        elif ph == "E":
            st = open_.get((e["name"], e.get("cat")))
            if st:
                add(e["name"], e["ts"] - st.pop())
    if format == "json":
        return json.dumps({k: {"count": v[0], "total_us": v[1], "min_us": v[2], "max_us": v[3]} for k, v in agg.items()})
    lines = ["%-40s %10s %14s %12s %12s %12s" % ("Name", "Count", "Total(us)", "Min(us)", "Max(us)", "Avg(us)")]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-40s %10d %14.1f %12.1f %12.1f %12.1f" % (k, v[0], v[1], v[2], v[3], v[1] / max(1, v[0])))
    return "\n".join(lines)


def dump(finished=True, profile_process="worker"):
    if profile_process == "server":
        _kv_handle.set_server_profiler_command(3, "1" if finished else "0"); return
    with _lock:
        evs = list(_events) + _resolve_device_ranges()
    meta = [{"name": "process_name", "ph": "M", "pid": "cpu", "args": {"name": "host"}}]
    with open(_cfg["filename"], "w") as f:
        json.dump({"traceEvents": meta + evs, "displayTimeUnit": "ms"}, f)
    if finished:
        _state["running"] = False


dump_profile = dump


class scope:
    """``with profiler.scope('name', device=True):`` → one complete ('X') event, host- or device-timed."""

    def __init__(self, name, cat="operator", device=False):
        self.name, self.cat, self.device = name, cat, device and torch.cuda.is_available()

    def __enter__(self):
        if not is_active():
            self._on = False; return self
        self._on = True
        self._ts = _now_us()
        if self.device:
            self._s = torch.cuda.Event(enable_timing=True); self._e = torch.cuda.Event(enable_timing=True)
            self._s.record()
        return self

    def __exit__(self, *a):
        if not self._on:
            return
        if self.device:
            self._e.record()
            _dev_ranges.append((self.name, self.cat, self._s, self._e, self._ts, torch.cuda.current_device()))
        else:
            _emit({"name": self.name, "cat": self.cat, "ph": "X", "ts": self._ts, "dur": _now_us() - self._ts,
                   "pid": "cpu", "tid": threading.get_ident() % 100000})


class Domain:
    def __init__(self, name):
        self.name = name

    def new_task(self, name): return Task(self, name)
    def new_frame(self, name): return Frame(self, name)
    def new_counter(self, name, value=None): return Counter(self, name, value)
    def new_marker(self, name): return Marker(self, name)

    def __str__(self):
        return self.name


class _Duration:
    _ph = ("B", "E")

    def __init__(self, domain, name):
        self.domain, self.name = domain, name

    def start(self):
        if is_active():
            _emit({"name": self.name, "cat": str(self.domain), "ph": "B", "ts": _now_us(), "pid": "cpu", "tid": 0})

    def stop(self):
        if is_active():
            _emit({"name": self.name, "cat": str(self.domain), "ph": "E", "ts": _now_us(), "pid": "cpu", "tid": 0})

    def __enter__(self):
        self.start(); return self

    def __exit__(self, *a):
        self.stop()

    def __str__(self):
        return self.name


class Task(_Duration):
    pass


class Frame(_Duration):
    pass


class Event(_Duration):
    def __init__(self, name):
        super().__init__("event", name)


class Counter:
    def __init__(self, domain, name, value=None):
        self.domain, self.name, self.value = domain, name, 0
        if value is not None:
            self.set_value(value)

    def set_value(self, value):
        self.value = value
        if is_active():
            _emit({"name": self.name, "cat": str(self.domain), "ph": "C", "ts": _now_us(), "pid": "cpu",
                   "args": {self.name: value}})

    def increment(self, delta=1): self.set_value(self.value + delta)
    def decrement(self, delta=1): self.set_value(self.value - delta)
    def __iadd__(self, v): self.increment(v); return self
    def __isub__(self, v): self.decrement(v); return self


class Marker:
    def __init__(self, domain, name):
        self.domain, self.name = domain, name

    def mark(self, scope="process"):
        if is_active():
            _emit({"name": self.name, "cat": str(self.domain), "ph": "i", "ts": _now_us(), "pid": "cpu", "tid": 0,
                   "s": {"global": "g", "process": "p", "thread": "t"}.get(scope, "p")})


# MXNET_PROFILER_AUTOSTART=1 starts profiling at import with MXNET_PROFILER_MODE (0: symbolic only, 1: everything) — profiler.cc:80-88
if os.environ.get("MXNET_PROFILER_AUTOSTART", "0") == "1":
    _all = os.environ.get("MXNET_PROFILER_MODE", "0") == "1"
    set_config(profile_all=_all, profile_symbolic=True, profile_imperative=_all, profile_api=_all, profile_memory=_all)
    set_state("run")
    import atexit as _atexit
    _atexit.register(lambda: dump(True) if _state["running"] else None)
