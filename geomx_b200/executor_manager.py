"""``mx.executor_manager`` — the pre-Module helper that drives one executor per device (parity: python/mxnet/executor_manager.py:
``DataParallelExecutorManager`` with ``install_monitor / set_params / copy_to / param_arrays / grad_arrays / aux_arrays / load_data_batch /
forward / backward / update_metric``).  Implemented as a thin shell over ``mx.mod.Module``, which owns the per-context executors and the
batch slicing; new code should use Module directly."""
from __future__ import annotations

import logging

from .module import Module

__all__ = ["DataParallelExecutorManager", "DataParallelExecutorGroup"]


def _split_input_slice(batch_size, work_load_list):
    """Slices of a batch proportional to ``work_load_list`` (executor_manager.py:32-70)."""
    total = float(sum(work_load_list))
    slices, start = [], 0
    for i, w in enumerate(work_load_list):
        end = batch_size if i == len(work_load_list) - 1 else min(batch_size, start + int(round(batch_size * w / total)))
        if end <= start:
            raise ValueError("Too many slices. Some splits are empty.")
        slices.append(slice(start, end)); start = end
    return slices


class DataParallelExecutorManager:
    def __init__(self, symbol, ctx, train_data, arg_names=None, param_names=None, aux_names=None, work_load_list=None, logger=None, sym_gen=None):
        self.logger = logger or logging
        self.symbol, self.ctx = symbol, (ctx if isinstance(ctx, (list, tuple)) else [ctx])
        data_names = [d[0] if isinstance(d, tuple) else d.name for d in train_data.provide_data]
        label_names = [d[0] if isinstance(d, tuple) else d.name for d in (train_data.provide_label or [])]
        label_names = [n for n in label_names if n in symbol.list_arguments()]
        self._mod = Module(symbol, data_names=data_names, label_names=label_names or None, context=list(self.ctx), logger=self.logger)
        self._mod.bind(train_data.provide_data, train_data.provide_label if label_names else None, for_training=True)
        self.param_names = param_names or self._mod._param_names
        self.aux_names = aux_names or self._mod._aux_names
        self.arg_names = arg_names or symbol.list_arguments()
        self.slices = self._mod._slices
        self._batch = None

    def install_monitor(self, monitor):
        self._mod.install_monitor(monitor)

    def set_params(self, arg_params, aux_params):
        self._mod.init_params(arg_params=arg_params, aux_params=aux_params, allow_missing=False, force_init=True)

    def copy_to(self, arg_params, aux_params):
        """Write the (device-averaged) current parameters into the given dicts."""
        arg, aux = self._mod.get_params()
        for k, v in arg.items():
            arg_params[k] = v.copy() if k not in arg_params else arg_params[k]
            arg_params[k][:] = v
        for k, v in aux.items():
            aux_params[k] = v.copy() if k not in aux_params else aux_params[k]
            aux_params[k][:] = v

    param_arrays = property(lambda self: self._mod._param_arrays())
    grad_arrays = property(lambda self: self._mod._grad_arrays())
    aux_arrays = property(lambda self: [[ex.aux_dict[n] for ex in self._mod._execs] for n in self.aux_names])

    def load_data_batch(self, data_batch):
        self._batch = data_batch

    def forward(self, is_train=False):
        self._mod.forward(self._batch, is_train=is_train)

    def backward(self):
        self._mod.backward()

    def update_metric(self, metric, labels, pre_sliced=False):
        self._mod.update_metric(metric, labels)


DataParallelExecutorGroup = DataParallelExecutorManager
