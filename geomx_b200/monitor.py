"""``mx.monitor.Monitor`` — periodic statistics of executor tensors (parity: ``python/mxnet/monitor.py``: install on executors, ``tic`` /
``toc`` / ``toc_print``; default statistic ``norm(x)/sqrt(size)``)."""
from __future__ import annotations

import logging
import re
from math import sqrt

__all__ = ["Monitor"]


class Monitor:
    def __init__(self, interval, stat_func=None, pattern=".*", sort=False):
        self.stat_func = stat_func or (lambda x: float(x.norm().asscalar()) / sqrt(max(1, x.size)))
        self.interval, self.activated, self.queue, self.step = interval, False, [], 0
        self.exes, self.re_prog, self.sort = [], re.compile(pattern), sort

    def install(self, exe):
        """Hooks the executor's per-node callback so intermediate outputs are sampled too (monitor.py:80-95)."""
        if hasattr(exe, "set_monitor_callback"):
            exe.set_monitor_callback(self.stat_helper)
        self.exes.append(exe)

    def stat_helper(self, name, array):
        if self.activated and self.re_prog.match(name):
            self.queue.append((self.step, name, self.stat_func(array)))

    def tic(self):
        if self.step % self.interval == 0:
            self.queue, self.activated = [], True
        self.step += 1

    def toc(self):
        if not self.activated:
            return []
        for exe in self.exes:
            for group in (exe.arg_dict, exe.aux_dict, exe.grad_dict):
                for name, arr in group.items():
                    if arr is not None and self.re_prog.match(name):
                        self.queue.append((self.step, name if group is not exe.grad_dict else name + "_grad", self.stat_func(arr)))
            for i, out in enumerate(exe.outputs):
                name = "output%d" % i
                if self.re_prog.match(name):
                    self.queue.append((self.step, name, self.stat_func(out)))
        self.activated = False
        res = sorted(self.queue, key=lambda x: x[1]) if self.sort else list(self.queue)
        self.queue = []
        return [(n, k, str(v)) for n, k, v in res]

    def toc_print(self):
        for n, k, v in self.toc():
            logging.info("Batch: %7d %30s %s", n, k, v)
