"""``mx.contrib.tensorboard.LogMetricsCallback`` — log evaluation metrics per batch / epoch (parity: python/mxnet/contrib/tensorboard.py:25-73).

With the ``tensorboard`` package installed the scalars go to a regular event file (``torch.utils.tensorboard.SummaryWriter``); without it
they are appended to ``<logging_dir>/scalars.jsonl`` (one ``{"tag", "value", "step", "wall_time"}`` object per line) so runs in an offline
image still leave a machine-readable trace."""
from __future__ import annotations

import json
import os
import time

__all__ = ["LogMetricsCallback", "JsonlSummaryWriter"]


class JsonlSummaryWriter:
    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self._f = open(os.path.join(logdir, "scalars.jsonl"), "a")

    def add_scalar(self, tag, value, global_step=None):
        self._f.write(json.dumps({"tag": tag, "value": float(value), "step": global_step, "wall_time": time.time()}) + "\n")
        self._f.flush()

    def close(self):
        self._f.close()


def _writer(logdir):
    try:
        import tensorboard  # noqa: F401
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(logdir)
    except Exception:  # noqa: BLE001  (package missing or unusable)
        return JsonlSummaryWriter(logdir)


class LogMetricsCallback:
    """Use as ``batch_end_callback`` / ``eval_end_callback`` of ``Module.fit``: ``LogMetricsCallback('logs/train', prefix='train')``."""

    def __init__(self, logging_dir, prefix=None):
        self.prefix, self.step = prefix, 0
        self.summary_writer = _writer(logging_dir)

    def __call__(self, param):
        if param.eval_metric is None:
            return
        self.step += 1
        for name, value in param.eval_metric.get_name_value():
            if self.prefix is not None:
                name = "%s-%s" % (self.prefix, name)
            self.summary_writer.add_scalar(name, value, self.step)
