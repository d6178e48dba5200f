"""Checkpoint helpers + kvstore glue of the symbolic/Module training path.

Parity: ``python/mxnet/model.py`` — ``save_checkpoint``/``load_checkpoint`` (:383-447: ``prefix-symbol.json`` +
``prefix-%04d.params`` with ``arg:``/``aux:`` name prefixes) and ``_create_kvstore`` / ``_initialize_kvstore`` /
``_update_params_on_kvstore`` / ``_update_params`` (:58-176) which drive the same KVStore from ``Module``."""
from __future__ import annotations

import json
import logging

from . import kvstore as kvs
from . import ndarray as nd

__all__ = ["save_checkpoint", "load_checkpoint", "load_params", "_create_kvstore", "_initialize_kvstore",
           "_update_params_on_kvstore", "_update_params", "BatchEndParam", "FeedForward"]

from collections import namedtuple

BatchEndParam = namedtuple("BatchEndParams", ["epoch", "nbatch", "eval_metric", "locals"])


def _create_kvstore(kvstore, num_device, arg_params):
    update_on_kvstore = True
    if kvstore is None:
        kv = None
    elif isinstance(kvstore, kvs.KVStoreBase):
        kv = kvstore
    elif isinstance(kvstore, str):
        if num_device == 1 and "dist" not in kvstore:
            kv = None
        else:
            kv = kvs.create(kvstore)
            if kvstore == "local":
                max_size = max((p.size for p in arg_params.values()), default=0)
                if max_size > 1024 * 1024 * 16:
                    update_on_kvstore = False
    else:
        raise TypeError("kvstore must be KVStore, str or None")
    if kv is None:
        update_on_kvstore = False
    return kv, update_on_kvstore


def _initialize_kvstore(kvstore, param_arrays, arg_params, param_names, update_on_kvstore):
    for idx, param_on_devs in enumerate(param_arrays):
        name = param_names[idx]
        kvstore.init(name, arg_params[name])
        if update_on_kvstore:
            kvstore.pull(name, param_on_devs, priority=-idx)


def _update_params_on_kvstore(param_arrays, grad_arrays, kvstore, param_names):
    for index, (arg_list, grad_list) in enumerate(zip(param_arrays, grad_arrays)):
        if grad_list[0] is None:
            continue
        name = param_names[index]
        kvstore.push(name, grad_list, priority=-index)
        kvstore.pull(name, arg_list, priority=-index)


def _update_params(param_arrays, grad_arrays, updater, num_device, kvstore=None, param_names=None):
    for i, (arg_list, grad_list) in enumerate(zip(param_arrays, grad_arrays)):
        if grad_list[0] is None:
            continue
        index = i
        if kvstore:
            name = param_names[index]
            kvstore.push(name, grad_list, priority=-index)
            kvstore.pull(name, grad_list, priority=-index)
        for k, (w, g) in enumerate(zip(arg_list, grad_list)):
            updater(index * num_device + k, g, w)


def save_checkpoint(prefix, epoch, symbol, arg_params, aux_params):
    if symbol is not None:
        with open("%s-symbol.json" % prefix, "w") as f:
            f.write(symbol if isinstance(symbol, str) else symbol.tojson() if hasattr(symbol, "tojson") else json.dumps(symbol))
    save_dict = {("arg:%s" % k): v for k, v in arg_params.items()}
    save_dict.update({("aux:%s" % k): v for k, v in (aux_params or {}).items()})
    param_name = "%s-%04d.params" % (prefix, epoch)
    nd.save(param_name, save_dict)
    logging.info('Saved checkpoint to "%s"', param_name)


def load_params(prefix, epoch):
    save_dict = nd.load("%s-%04d.params" % (prefix, epoch))
    arg_params, aux_params = {}, {}
    for k, v in save_dict.items():
        tp, name = k.split(":", 1)
        if tp == "arg":
            arg_params[name] = v
        elif tp == "aux":
            aux_params[name] = v
    return arg_params, aux_params


def load_checkpoint(prefix, epoch):
    """``(symbol, arg_params, aux_params)`` of ``prefix-symbol.json`` + ``prefix-%04d.params`` (python/mxnet/model.py:414-450).  The symbol is
    a ``Symbol`` (either JSON dialect); ``None`` when the file is absent or holds no graph."""
    from . import symbol as sym
    symbol = None
    try:
        with open("%s-symbol.json" % prefix) as f:
            text = f.read()
        symbol = sym.load_json(text) if '"nodes"' in text else None
    except FileNotFoundError:
        pass
    arg_params, aux_params = load_params(prefix, epoch)
    return symbol, arg_params, aux_params


class FeedForward:
    """The pre-Module training API, kept as a thin shell over ``mx.mod.Module`` (parity: python/mxnet/model.py FeedForward :450-1000:
    ``fit`` / ``predict`` / ``score`` / ``save`` / ``load`` / ``create``).  ``X`` may be a DataIter or a numpy / NDArray matrix with ``y``."""

    def __init__(self, symbol, ctx=None, num_epoch=None, epoch_size=None, optimizer="sgd", initializer=None, numpy_batch_size=128,
                 arg_params=None, aux_params=None, allow_extra_params=False, begin_epoch=0, **kwargs):
        self.symbol, self.ctx, self.num_epoch, self.optimizer, self.initializer = symbol, ctx, num_epoch, optimizer, initializer
        self.numpy_batch_size, self.arg_params, self.aux_params, self.begin_epoch, self.kwargs = numpy_batch_size, arg_params, aux_params, begin_epoch, dict(kwargs)
        self._mod = None

    def _iter(self, X, y=None, shuffle=False):
        from . import io
        if isinstance(X, io.DataIter):
            return X
        import numpy as _np
        if y is None:
            y = _np.zeros((len(X),), dtype=_np.float32)
        return io.NDArrayIter(X, y, batch_size=min(self.numpy_batch_size, len(X)), shuffle=shuffle, last_batch_handle="pad")

    def _module(self, data):
        from .module import Module
        if self._mod is None:
            label_names = [d[0] if isinstance(d, tuple) else d.name for d in (data.provide_label or [])]
            label_names = [n for n in label_names if n in self.symbol.list_arguments()]
            self._mod = Module(self.symbol, data_names=[d[0] if isinstance(d, tuple) else d.name for d in data.provide_data],
                               label_names=label_names or None, context=self.ctx)
        return self._mod

    def fit(self, X, y=None, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None, kvstore="local", logger=None,
            work_load_list=None, monitor=None, eval_end_callback=None, eval_batch_end_callback=None):
        data = self._iter(X, y, shuffle=True)
        if isinstance(eval_data, tuple):
            eval_data = self._iter(*eval_data)
        mod = self._module(data)
        opt_params = dict(self.kwargs) or {"learning_rate": 0.01}
        mod.fit(data, eval_data=eval_data, eval_metric=eval_metric, epoch_end_callback=epoch_end_callback, batch_end_callback=batch_end_callback,
                kvstore=kvstore, optimizer=self.optimizer, optimizer_params=opt_params, initializer=self.initializer, arg_params=self.arg_params,
                aux_params=self.aux_params, allow_missing=self.arg_params is not None, begin_epoch=self.begin_epoch, num_epoch=self.num_epoch)
        self.arg_params, self.aux_params = mod.get_params()
        return self

    def _bound(self, data):
        mod = self._module(data)
        if not mod.binded:
            mod.bind(data.provide_data, data.provide_label, for_training=False)
            mod.init_params(arg_params=self.arg_params, aux_params=self.aux_params, allow_missing=False)
        return mod

    def predict(self, X, num_batch=None, return_data=False, reset=True):
        data = self._iter(X)
        out = self._bound(data).predict(data, num_batch=num_batch, reset=reset)
        return out.asnumpy() if hasattr(out, "asnumpy") else [o.asnumpy() for o in out]

    def score(self, X, eval_metric="acc", num_batch=None, batch_end_callback=None, reset=True):
        data = self._iter(X)
        return self._bound(data).score(data, eval_metric, num_batch=num_batch, reset=reset)[0][1]

    def save(self, prefix, epoch=None):
        save_checkpoint(prefix, self.num_epoch if epoch is None else epoch, self.symbol.tojson(), self.arg_params, self.aux_params or {})

    @staticmethod
    def load(prefix, epoch, ctx=None, **kwargs):
        from . import symbol as sym
        net, arg, aux = load_checkpoint(prefix, epoch)
        return FeedForward(net, ctx=ctx, arg_params=arg, aux_params=aux, begin_epoch=epoch, **kwargs)

    @staticmethod
    def create(symbol, X, y=None, ctx=None, num_epoch=None, epoch_size=None, optimizer="sgd", initializer=None, eval_data=None, eval_metric="acc",
               epoch_end_callback=None, batch_end_callback=None, kvstore="local", logger=None, work_load_list=None, **kwargs):
        model = FeedForward(symbol, ctx=ctx, num_epoch=num_epoch, epoch_size=epoch_size, optimizer=optimizer, initializer=initializer, **kwargs)
        return model.fit(X, y, eval_data=eval_data, eval_metric=eval_metric, epoch_end_callback=epoch_end_callback,
                         batch_end_callback=batch_end_callback, kvstore=kvstore)
