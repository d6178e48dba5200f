"""``gluon.Trainer``.

Parity: ``python/mxnet/gluon/trainer.py:27-434`` — constructor (params dict/list, optimizer str/obj,
``kvstore`` str/obj/None, ``update_on_kvstore``, ``compression_params``), ``_init_kvstore`` decision table
(:169-246), ``step`` :258 (``rescale_grad = scale / batch_size``), ``allreduce_grads`` :301, ``update`` :333,
``save_states`` :387 / ``load_states`` :410 (pickled ``Updater`` states, incl. optimizer), ``learning_rate`` /
``set_learning_rate``.

B200 design: on CUDA with a native-spec optimizer (SGD / momentum / Adam / DCASGD) and no kvstore the whole
parameter set is updated by ONE multi-tensor fused kernel launch (``ops.native.multi_tensor_*``) instead of
one launch per parameter (the reference issues 3 mshadow passes per Adam parameter).
"""
from __future__ import annotations

import torch

from .. import optimizer as opt
from .parameter import Parameter, ParameterDict

__all__ = ["Trainer"]


class Trainer:
    def __init__(self, params, optimizer, optimizer_params=None, kvstore="device", compression_params=None,
                 update_on_kvstore=None):
        if isinstance(params, (dict, ParameterDict)):
            params = list(params.values())
        if not isinstance(params, (list, tuple)):
            raise ValueError("First argument must be a list or dict of Parameters, got %s." % type(params))
        self._params, self._param2idx = [], {}
        for i, p in enumerate(params):
            if not isinstance(p, Parameter):
                raise ValueError("First argument must be a list or dict of Parameters, got list of %s." % type(p))
            self._param2idx[p.name] = i
            self._params.append(p)
            p._trainer = self
        self._compression_params = compression_params
        optimizer_params = optimizer_params or {}
        self._scale = float(optimizer_params.get("rescale_grad", 1.0))
        self._contexts = self._check_contexts()
        self._init_optimizer(optimizer, optimizer_params)
        self._kvstore_params = {"kvstore": kvstore, "update_on_kvstore": update_on_kvstore}
        self._kv_initialized = False
        self._kvstore = None
        self._update_on_kvstore = None
        self._params_to_init = []
        self._reset_kvstore()

    def _check_contexts(self):
        contexts = None
        for p in self._params:
            ctx = p.list_ctx()
            assert contexts is None or contexts == ctx, \
                "All Parameters must be initialized on the same set of contexts, but Parameter %s is initialized on %s while previous Parameters are initialized on %s." % (p.name, str(ctx), str(contexts))
            contexts = ctx
        return contexts

    def _init_optimizer(self, optimizer, optimizer_params):
        param_dict = {i: p for i, p in enumerate(self._params)}
        if isinstance(optimizer, opt.Optimizer):
            assert not optimizer_params, "optimizer_params must be None if optimizer is an instance of Optimizer instead of str"
            self._optimizer = optimizer
            self._optimizer.param_dict = param_dict
        else:
            self._optimizer = opt.create(optimizer, param_dict=param_dict, **optimizer_params)
        self._updaters = [opt.get_updater(self._optimizer) for _ in self._contexts]

    def _reset_kvstore(self):
        if self._kvstore and "dist" in self._kvstore.type:
            raise RuntimeError("Cannot reset distributed KVStore.")
        self._kv_initialized = False
        self._kvstore = None
        self._update_on_kvstore = None
        self._params_to_init = [p for p in self._params]

    def _init_kvstore(self):
        from .. import kvstore as kvs
        config = self._kvstore_params
        kv, update_on_kv = config["kvstore"], config["update_on_kvstore"]
        if kv is None or (isinstance(kv, str) and len(self._contexts) == 1 and "dist" not in kv):
            kvstore, update_on_kvstore = None, False
        else:
            kvstore = kvs.create(kv) if isinstance(kv, str) else kv
            if update_on_kv is None:
                update_on_kvstore = "dist" in kvstore.type
            else:
                update_on_kvstore = bool(update_on_kv)
            if "dist" in kvstore.type and "async" in kvstore.type and not update_on_kvstore:
                raise ValueError("Please set update_on_kvstore to true when training in async mode.")
        if kvstore:
            if self._compression_params:
                kvstore.set_gradient_compression(self._compression_params)
            if update_on_kvstore:
                kvstore.set_optimizer(self._optimizer)
            self._kvstore, self._update_on_kvstore = kvstore, update_on_kvstore
        else:
            self._kvstore, self._update_on_kvstore = None, None
        self._kv_initialized = True

    def _init_params(self):
        assert self._kv_initialized
        if not self._kvstore:
            self._params_to_init = []
            return
        rest = []
        for p in self._params_to_init:
            if p._deferred_init:
                rest.append(p)
            else:
                idx = self._param2idx[p.name]
                self._kvstore.init(idx, p.list_data()[0])
                if self._update_on_kvstore or True:
                    self._kvstore.pull(idx, p.list_data(), priority=-idx)
        self._params_to_init = rest

    @property
    def learning_rate(self):
        return self._optimizer.learning_rate

    @property
    def optimizer(self):
        return self._optimizer

    def set_learning_rate(self, lr):
        self._optimizer.set_learning_rate(lr)

    def step(self, batch_size, ignore_stale_grad=False):
        rescale_grad = self._scale / batch_size
        self._check_and_rescale_grad(rescale_grad)
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        self._allreduce_grads()
        self._update(ignore_stale_grad)

    def _check_and_rescale_grad(self, scale):
        if self._update_on_kvstore and self._kv_initialized and self._kvstore and "dist" in self._kvstore.type:
            if self._optimizer.rescale_grad != scale:
                raise UserWarning("Possible change in the `batch_size` from previous `step` detected. "
                                  "Optimizer gradient normalizing factor will not change w.r.t new batch_size when "
                                  "update_on_kvstore=True and when distributed kvstore is used.")
        self._optimizer.rescale_grad = scale

    def allreduce_grads(self):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        assert not (self._kvstore and self._update_on_kvstore), \
            "allreduce_grads() when parameters are updated on kvstore is not supported. Try setting `update_on_kvstore` to False when creating trainer."
        self._allreduce_grads()

    def _allreduce_grads(self):
        if not self._kvstore:
            return
        for i, p in enumerate(self._params):
            if p.grad_req != "null":
                self._kvstore.push(i, p.list_grad(), priority=-i)
                if not self._update_on_kvstore:
                    self._kvstore.pull(i, p.list_grad(), priority=-i)

    def update(self, batch_size, ignore_stale_grad=False):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        assert not (self._kvstore and self._update_on_kvstore), \
            "update() when parameters are updated on kvstore is not supported. Try setting `update_on_kvstore` to False when creating trainer."
        self._check_and_rescale_grad(self._scale / batch_size)
        self._update(ignore_stale_grad)

    def _update(self, ignore_stale_grad=False):
        if self._kvstore and self._update_on_kvstore:
            for i, p in enumerate(self._params):
                if p.grad_req != "null":
                    self._kvstore.pull(i, p.list_data(), priority=-i)
            return
        if self._try_fused_update():
            return
        for i, p in enumerate(self._params):
            if p.grad_req == "null":
                continue
            for upd, arr, grad in zip(self._updaters, p.list_data(), p.list_grad()):
                upd(i, grad, arr)

    def _try_fused_update(self):
        """One multi-tensor launch for the whole parameter set when a native spec exists (CUDA, fp32, 1 ctx)."""
        from ..ops import native
        if len(self._contexts) != 1 or not native.available():
            return False
        spec = self._optimizer.spec()
        if spec is None or spec["name"] not in ("adam", "sgd") or self._optimizer.lr_scheduler is not None:
            return False
        ws, gs, idxs = [], [], []
        for i, p in enumerate(self._params):
            if p.grad_req == "null":
                continue
            w = p.list_data()[0]._t; g = p.list_grad()[0]._t
            if not (w.is_cuda and w.dtype == torch.float32 and g.dtype == torch.float32):
                return False
            ws.append(w.detach()); gs.append(g); idxs.append(i)
        if not ws:
            return True
        return native.multi_tensor_update(self._optimizer, self._updaters[0], idxs, ws, gs)

    def save_states(self, fname):
        assert self._optimizer is not None
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        if self._update_on_kvstore:
            assert not self._params_to_init, "Cannot save trainer states when some parameters are not yet initialized in kvstore."
            self._kvstore.save_optimizer_states(fname, dump_optimizer=True)
        else:
            with open(fname, "wb") as f:
                f.write(self._updaters[0].get_states(dump_optimizer=True))

    def load_states(self, fname):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        if self._update_on_kvstore:
            self._kvstore.load_optimizer_states(fname)
            self._optimizer = self._kvstore._updater.optimizer
        else:
            with open(fname, "rb") as f:
                states = f.read()
            for upd in self._updaters:
                upd.set_states(states)
                upd.optimizer = self._updaters[0].optimizer
            self._optimizer = self._updaters[0].optimizer
        param_dict = {i: p for i, p in enumerate(self._params)}
        self._optimizer.param_dict = param_dict
