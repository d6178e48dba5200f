"""Training callbacks (``mx.callback``): checkpointing, metric logging, throughput and progress reporting.

The callback *protocol* is the reference's (``python/mxnet/callback.py`` / ``module/base_module.py``): epoch-end callbacks are called as
``cb(epoch, symbol, arg_params, aux_params)``, batch-end callbacks as ``cb(BatchEndParam)`` where the param carries ``epoch``, ``nbatch`` and
``eval_metric``.  The implementations are built on two small pieces of this module: :class:`Every` (a period gate) and
:class:`ThroughputWindow` (a monotonic-clock sample counter), so that "every N batches" and "samples per second" mean one thing everywhere.
"""
from __future__ import annotations

import logging
import sys
import time

from .model import save_checkpoint

__all__ = ["do_checkpoint", "module_checkpoint", "log_train_metric", "Speedometer", "ProgressBar", "LogValidationMetricsCallback", "Every",
           "ThroughputWindow"]

_log = logging.getLogger(__name__)


class Every:
    """Period gate: ``gate(i)`` is true for the ``period``-th, ``2*period``-th, ... event when events are numbered from ``first`` (0 or 1)."""

    def __init__(self, period, first=0):
        self.period, self.first = max(1, int(period)), int(first)

    def __call__(self, index):
        return (int(index) - self.first + 1) % self.period == 0


class ThroughputWindow:
    """Samples per second over the window since the last :meth:`restart` (``time.perf_counter``: monotonic, unaffected by clock changes)."""

    def __init__(self):
        self.restart()

    def restart(self):
        self.t0, self.samples = time.perf_counter(), 0

    def add(self, n):
        self.samples += n

    def rate(self):
        return self.samples / max(time.perf_counter() - self.t0, 1e-9)


def _metric_pairs(param):
    metric = getattr(param, "eval_metric", None)
    return [] if metric is None else list(metric.get_name_value())


def module_checkpoint(mod, prefix, period=1, save_optimizer_states=False):
    """Epoch-end callback that saves ``mod`` (symbol, parameters, optionally optimizer states) after every ``period`` epochs."""
    due = Every(period)

    def on_epoch_end(epoch, symbol=None, arg_params=None, aux_params=None):
        if due(epoch):
            mod.save_checkpoint(prefix, epoch + 1, save_optimizer_states)
    return on_epoch_end


def do_checkpoint(prefix, period=1):
    """Epoch-end callback writing ``prefix-symbol.json`` / ``prefix-%04d.params`` after every ``period`` epochs."""
    due = Every(period)

    def on_epoch_end(epoch, symbol, arg_params, aux_params):
        if due(epoch):
            save_checkpoint(prefix, epoch + 1, symbol.tojson() if hasattr(symbol, "tojson") else symbol, arg_params, aux_params)
    return on_epoch_end


def log_train_metric(period, auto_reset=False):
    """Batch-end callback logging the running training metric every ``period`` batches (batch 0 included)."""
    period = max(1, int(period))

    def on_batch_end(param):
        if param.nbatch % period:
            return
        pairs = _metric_pairs(param)
        for name, value in pairs:
            _log.info("Iter[%d] Batch[%d] Train-%s=%f", param.epoch, param.nbatch, name, value)
        if pairs and auto_reset:
            param.eval_metric.reset()
    return on_batch_end


class Speedometer:
    """Batch-end callback: every ``frequent`` batches log the throughput of the window since the previous report (and the running metric,
    which is reset afterwards when ``auto_reset``).  ``last_speed`` keeps the most recent figure for programmatic use."""

    def __init__(self, batch_size, frequent=50, auto_reset=True):
        self.batch_size, self.frequent, self.auto_reset = int(batch_size), max(1, int(frequent)), auto_reset
        self.last_speed = None
        self._window = None          # opened by the first batch of an epoch
        self._prev_batch = -1

    def __call__(self, param):
        nbatch = param.nbatch
        if self._window is None or nbatch < self._prev_batch:      # first call, or the batch counter restarted (new epoch)
            self._window = ThroughputWindow()
            self._prev_batch = nbatch
            return
        self._window.add((nbatch - self._prev_batch) * self.batch_size)
        self._prev_batch = nbatch
        if nbatch % self.frequent:
            return
        self.last_speed = self._window.rate()
        pairs = _metric_pairs(param)
        text = "".join("\t%s=%f" % (n, v) for n, v in pairs)
        _log.info("Epoch[%d] Batch [%d]\tSpeed: %.2f samples/sec%s", param.epoch, nbatch, self.last_speed, text)
        if pairs and self.auto_reset:
            param.eval_metric.reset()
        self._window.restart()


class ProgressBar:
    """Batch-end callback drawing a text progress bar for an epoch of ``total`` batches."""

    def __init__(self, total, length=80):
        self.total, self.length = max(1, int(total)), int(length)

    def __call__(self, param):
        done = min(max(param.nbatch / self.total, 0.0), 1.0)
        filled = int(round(self.length * done))
        sys.stdout.write("[%s%s] %d%%\r" % ("=" * filled, "-" * (self.length - filled), int(-(-100.0 * done // 1))))
        sys.stdout.flush()


class LogValidationMetricsCallback:
    """Eval-end callback logging every validation metric of the epoch."""

    def __call__(self, param):
        for name, value in _metric_pairs(param):
            _log.info("Epoch[%d] Validation-%s=%f", param.epoch, name, value)
