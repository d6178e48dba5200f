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
           "_update_params_on_kvstore", "_update_params", "BatchEndParam"]

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
            f.write(symbol if isinstance(symbol, str) else json.dumps(symbol))
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
    symbol = None
    try:
        with open("%s-symbol.json" % prefix) as f:
            symbol = f.read()
    except FileNotFoundError:
        pass
    arg_params, aux_params = load_params(prefix, epoch)
    return symbol, arg_params, aux_params
