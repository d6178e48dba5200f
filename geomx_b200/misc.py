"""``mx.misc`` — the pre-``lr_scheduler`` learning-rate helpers kept for old scripts (parity: python/mxnet/misc.py)."""
from __future__ import annotations

import logging
import math


class LearningRateScheduler:
    """Base class: ``__call__(iteration) -> lr`` starting from ``base_lr``."""

    def __init__(self):
        self.base_lr = 0.01

    def __call__(self, iteration):
        raise NotImplementedError("must override this")


class FactorScheduler(LearningRateScheduler):
    """``lr = base_lr * factor^(iteration // step)``; logs when the rate changes."""

    def __init__(self, step, factor=0.1):
        super().__init__()
        if step < 1:
            raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor >= 1.0:
            raise ValueError("Factor must be less than 1 to make lr reduce")
        self.step, self.factor, self.old_lr, self.init = step, factor, self.base_lr, False

    def __call__(self, iteration):
        if not self.init:
            self.init, self.old_lr = True, self.base_lr
        lr = self.base_lr * math.pow(self.factor, int(iteration / self.step))
        if lr != self.old_lr:
            self.old_lr = lr
            logging.info("At Iteration [%d]: Swith to new learning rate %.5f", iteration, lr)
        return lr
