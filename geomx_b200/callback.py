"""Training callbacks (``mx.callback``).  Parity: ``python/mxnet/callback.py`` — ``do_checkpoint`` :55-85, ``module_checkpoint`` :27-52,
``log_train_metric`` :88-114, ``Speedometer`` :117-175, ``ProgressBar`` :178-207."""
from __future__ import annotations

import logging
import math
import sys
import time

from .model import save_checkpoint

__all__ = ["do_checkpoint", "module_checkpoint", "log_train_metric", "Speedometer", "ProgressBar", "LogValidationMetricsCallback"]


def module_checkpoint(mod, prefix, period=1, save_optimizer_states=False):
    period = int(max(1, period))

    def _callback(iter_no, sym=None, arg=None, aux=None):
        if (iter_no + 1) % period == 0:
            mod.save_checkpoint(prefix, iter_no + 1, save_optimizer_states)
    return _callback


def do_checkpoint(prefix, period=1):
    period = int(max(1, period))

    def _callback(iter_no, sym, arg, aux):
        if (iter_no + 1) % period == 0:
            save_checkpoint(prefix, iter_no + 1, sym.tojson() if hasattr(sym, "tojson") else sym, arg, aux)
    return _callback


def log_train_metric(period, auto_reset=False):
    def _callback(param):
        if param.nbatch % period == 0 and param.eval_metric is not None:
            for name, value in param.eval_metric.get_name_value():
                logging.info("Iter[%d] Batch[%d] Train-%s=%f", param.epoch, param.nbatch, name, value)
            if auto_reset:
                param.eval_metric.reset()
    return _callback


class Speedometer:
    """Logs samples/sec (and the running metric) every ``frequent`` batches."""

    def __init__(self, batch_size, frequent=50, auto_reset=True):
        self.batch_size, self.frequent, self.auto_reset = batch_size, frequent, auto_reset
        self.init, self.tic, self.last_count = False, 0.0, 0
        self.last_speed = None

    def __call__(self, param):
        count = param.nbatch
        if self.last_count > count:
            self.init = False
        self.last_count = count
        if self.init:
            if count % self.frequent == 0:
                speed = self.frequent * self.batch_size / max(time.time() - self.tic, 1e-9)
                self.last_speed = speed
                if param.eval_metric is not None:
                    nv = param.eval_metric.get_name_value()
                    if self.auto_reset:
                        param.eval_metric.reset()
                    msg = "Epoch[%d] Batch [%d]\tSpeed: %.2f samples/sec" + "\t%s=%f" * len(nv)
                    logging.info(msg, param.epoch, count, speed, *sum(nv, ()))
                else:
                    logging.info("Iter[%d] Batch [%d]\tSpeed: %.2f samples/sec", param.epoch, count, speed)
                self.tic = time.time()
        else:
            self.init = True
            self.tic = time.time()


class ProgressBar:
    def __init__(self, total, length=80):
        self.bar_len, self.total = length, total

    def __call__(self, param):
        filled = int(round(self.bar_len * param.nbatch / float(self.total)))
        pct = math.ceil(100.0 * param.nbatch / float(self.total))
        sys.stdout.write("[%s] %s%s\r" % ("=" * filled + "-" * (self.bar_len - filled), pct, "%"))


class LogValidationMetricsCallback:
    def __call__(self, param):
        if not param.eval_metric:
            return
        for name, value in param.eval_metric.get_name_value():
            logging.info("Epoch[%d] Validation-%s=%f", param.epoch, name, value)
