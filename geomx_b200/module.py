"""``mx.mod.Module`` — the symbolic training front end (bind → init_params → init_optimizer → fit / forward / backward / update).

Parity: ``python/mxnet/module/module.py`` (``bind`` :363-470, ``init_params`` :265-340, ``init_optimizer`` :472-560 incl. the kvstore setup
:501-543, ``forward``/``backward``/``update`` :588-668, ``save_checkpoint`` / ``load``), ``module/base_module.py`` (``fit`` :376-560, ``score``,
``predict``) and ``module/executor_group.py`` (``DataParallelExecutorGroup``: one executor per context, the batch split along axis 0).
The kvstore glue is ``model.py``'s (``_create_kvstore`` / ``_initialize_kvstore`` / ``_update_params(_on_kvstore)``), i.e. a Module trains
through the same KVStore implementations — local, device, the TCP HiPS client or the NVSwitch fabric — as the Gluon scripts do."""
from __future__ import annotations

import logging
import time

from . import initializer as init_mod
from . import metric as metric_mod
from . import ndarray as nd
from . import optimizer as opt
from .base import MXNetError
from .context import cpu
from .io import DataBatch, DataDesc
from .model import BatchEndParam, _create_kvstore, _initialize_kvstore, _update_params, _update_params_on_kvstore, load_checkpoint, save_checkpoint

__all__ = ["Module", "BaseModule", "BucketingModule", "SequentialModule", "PythonModule", "PythonLossModule"]


def _as_desc(shapes):
    out = []
    for s in shapes or []:
        out.append(s if isinstance(s, DataDesc) else DataDesc(s[0], tuple(s[1])))
    return out


class BaseModule:
    def __init__(self, logger=logging):
        self.logger = logger
        self.binded = self.for_training = self.params_initialized = self.optimizer_initialized = False

    # ---- what a concrete module provides (base_module.py:840-1050: the computation / parameter / optimizer interface)
    def _abstract(self, what):
        raise NotImplementedError("%s must be implemented by %s" % (what, type(self).__name__))

    data_names = property(lambda self: self._abstract("data_names"))
    output_names = property(lambda self: self._abstract("output_names"))
    data_shapes = property(lambda self: self._abstract("data_shapes"))
    label_shapes = property(lambda self: self._abstract("label_shapes"))
    output_shapes = property(lambda self: self._abstract("output_shapes"))
    symbol = property(lambda self: self._abstract("symbol"))

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False, shared_module=None, grad_req="write"):
        self._abstract("bind")

    def init_params(self, initializer=None, arg_params=None, aux_params=None, allow_missing=False, force_init=False, allow_extra=False):
        self._abstract("init_params")

    def get_params(self):
        self._abstract("get_params")

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        self.init_params(initializer=None, arg_params=arg_params, aux_params=aux_params, allow_missing=allow_missing, force_init=force_init,
                         allow_extra=allow_extra)

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), force_init=False):
        self._abstract("init_optimizer")

    def forward(self, data_batch, is_train=None):
        self._abstract("forward")

    def backward(self, out_grads=None):
        self._abstract("backward")

    def update(self):
        self._abstract("update")

    def get_outputs(self, merge_multi_context=True):
        self._abstract("get_outputs")

    def get_input_grads(self, merge_multi_context=True):
        self._abstract("get_input_grads")

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        self._abstract("update_metric")

    # ---- high level API shared by every module type
    def iter_predict(self, eval_data, num_batch=None, reset=True):
        """Generator of ``(outputs, batch index, batch)`` with the padding rows removed (base_module.py:290-330)."""
        assert self.binded and self.params_initialized
        if reset:
            eval_data.reset()
        for nbatch, batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.forward(batch, is_train=False)
            pad = batch.pad or 0
            yield [o[0:o.shape[0] - pad] for o in self.get_outputs()], nbatch, batch

    def save_params(self, fname):
        arg, aux = self.get_params()
        d = {"arg:%s" % k: v.as_in_context(cpu()) for k, v in arg.items()}
        d.update({"aux:%s" % k: v.as_in_context(cpu()) for k, v in aux.items()})
        nd.save(fname, d)

    def load_params(self, fname):
        arg, aux = {}, {}
        for k, v in nd.load(fname).items():
            kind, name = k.split(":", 1)
            if kind == "arg":
                arg[name] = v
            elif kind == "aux":
                aux[name] = v
            else:
                raise ValueError("Invalid param file " + fname)
        self.set_params(arg, aux)

    def install_monitor(self, mon):
        raise NotImplementedError

    def prepare(self, data_batch, sparse_row_id_fn=None):
        """Hook called before ``forward`` (bucket switch, row_sparse pulls); nothing to do for dense single-graph modules."""

    def get_states(self, merge_multi_context=True):
        return []

    def set_states(self, states=None, value=None):
        assert not states and not value

    def forward_backward(self, data_batch):
        self.forward(data_batch, is_train=True)
        self.backward()

    def score(self, eval_data, eval_metric, num_batch=None, reset=True, epoch=0):
        assert self.binded and self.params_initialized
        if reset:
            eval_data.reset()
        eval_metric = metric_mod.create(eval_metric) if not isinstance(eval_metric, metric_mod.EvalMetric) else eval_metric
        eval_metric.reset()
        for nbatch, batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.forward(batch, is_train=False)
            self.update_metric(eval_metric, batch.label)
        return eval_metric.get_name_value()

    def predict(self, eval_data, num_batch=None, merge_batches=True, reset=True):
        assert self.binded and self.params_initialized
        if reset:
            eval_data.reset()
        outs = []
        for nbatch, batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.forward(batch, is_train=False)
            pad = batch.pad or 0
            outs.append([o[0:o.shape[0] - pad].copy() for o in self.get_outputs()])
        if not outs:
            return outs
        if merge_batches:
            merged = [nd.concat(*[b[i] for b in outs], dim=0) for i in range(len(outs[0]))]
            return merged[0] if len(merged) == 1 else merged
        return outs

    def fit(self, train_data, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None, kvstore="local",
            optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), eval_end_callback=None, initializer=None, arg_params=None,
            aux_params=None, allow_missing=False, force_rebind=False, force_init=False, begin_epoch=0, num_epoch=None, validation_metric=None):
        assert num_epoch is not None, "please specify number of epochs"
        self.bind(data_shapes=train_data.provide_data, label_shapes=train_data.provide_label, for_training=True, force_rebind=force_rebind)
        self.init_params(initializer=initializer or init_mod.Uniform(0.01), arg_params=arg_params, aux_params=aux_params,
                         allow_missing=allow_missing, force_init=force_init)
        self.init_optimizer(kvstore=kvstore, optimizer=optimizer, optimizer_params=optimizer_params)
        eval_metric = metric_mod.create(eval_metric) if not isinstance(eval_metric, metric_mod.EvalMetric) else eval_metric
        validation_metric = validation_metric or eval_metric
        for epoch in range(begin_epoch, num_epoch):
            tic = time.time()
            eval_metric.reset()
            train_data.reset()
            for nbatch, batch in enumerate(train_data):
                self.forward_backward(batch)
                self.update()
                self.update_metric(eval_metric, batch.label)
                if batch_end_callback is not None:
                    p = BatchEndParam(epoch=epoch, nbatch=nbatch, eval_metric=eval_metric, locals=locals())
                    for cb in (batch_end_callback if isinstance(batch_end_callback, (list, tuple)) else [batch_end_callback]):
                        cb(p)
            for name, val in eval_metric.get_name_value():
                self.logger.info("Epoch[%d] Train-%s=%f", epoch, name, val)
            self.logger.info("Epoch[%d] Time cost=%.3f", epoch, time.time() - tic)
            if epoch_end_callback is not None:
                arg, aux = self.get_params()
                for cb in (epoch_end_callback if isinstance(epoch_end_callback, (list, tuple)) else [epoch_end_callback]):
                    cb(epoch, self.symbol, arg, aux)
            if eval_data is not None:
                for name, val in self.score(eval_data, validation_metric, epoch=epoch):
                    self.logger.info("Epoch[%d] Validation-%s=%f", epoch, name, val)


class Module(BaseModule):
    def __init__(self, symbol, data_names=("data",), label_names=("softmax_label",), logger=logging, context=None, fixed_param_names=None):
        super().__init__(logger)
        self._symbol = symbol
        ctx = context if context is not None else cpu()
        self._context = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx]
        self._data_names, self._label_names = list(data_names or []), list(label_names or [])
        args = symbol.list_arguments()
        self._param_names = [a for a in args if a not in self._data_names + self._label_names]
        self._fixed = set(fixed_param_names or [])
        self._aux_names = symbol.list_auxiliary_states()
        self._execs, self._arg_params, self._aux_params = [], None, None
        self._kvstore, self._update_on_kvstore, self._updater, self._optimizer = None, False, None, None
        self._slices = []

    symbol = property(lambda self: self._symbol)
    data_names = property(lambda self: self._data_names)
    output_names = property(lambda self: self._symbol.list_outputs())

    # ---- bind: one executor per context, batch split along axis 0 (DataParallelExecutorGroup.decide_slices)
    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False, shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        self.for_training, self._inputs_need_grad = for_training, inputs_need_grad
        self._data_shapes, self._label_shapes = _as_desc(data_shapes), _as_desc(label_shapes)
        batch = self._data_shapes[0].shape[0]
        n = len(self._context)
        bounds = [(batch * i) // n for i in range(n + 1)]
        self._slices = [slice(bounds[i], bounds[i + 1]) for i in range(n)]
        self._execs = []
        for ctx, sl in zip(self._context, self._slices):
            shapes = {d.name: (sl.stop - sl.start,) + tuple(d.shape[1:]) for d in self._data_shapes + self._label_shapes}
            req = {}
            for a in self._symbol.list_arguments():
                if a in self._param_names:
                    req[a] = "null" if (not for_training or a in self._fixed) else grad_req
                elif a in self._data_names:
                    req[a] = grad_req if inputs_need_grad else "null"
                else:
                    req[a] = "null"
            self._execs.append(self._symbol.simple_bind(ctx, grad_req=req, **shapes))
        self.binded = True
        if shared_module is not None:
            # bucketing: parameters, gradients and aux states are the SAME arrays as the shared module's (module.py:420-440 shares the memory pool)
            assert shared_module.binded and shared_module.params_initialized and len(shared_module._execs) == len(self._execs)
            for ex, sex in zip(self._execs, shared_module._execs):
                for n in self._param_names:
                    if n in sex.arg_dict:
                        ex.arg_dict[n] = sex.arg_dict[n]
                        if n in sex.grad_dict and n in ex.grad_dict:
                            ex.grad_dict[n] = sex.grad_dict[n]
                for n in self._aux_names:
                    if n in sex.aux_dict:
                        ex.aux_dict[n] = sex.aux_dict[n]
                ex.arg_arrays = [ex.arg_dict[n] for n in self._symbol.list_arguments()]
                ex.grad_arrays = [ex.grad_dict.get(n) for n in self._symbol.list_arguments()]
                ex.aux_arrays = [ex.aux_dict[n] for n in self._aux_names]
            self._arg_params, self._aux_params, self.params_initialized = shared_module._arg_params, shared_module._aux_params, True
        elif self._arg_params is not None:
            self._sync_params_to_devices()

    # ---- parameters
    def init_params(self, initializer=None, arg_params=None, aux_params=None, allow_missing=False, force_init=False, allow_extra=False):
        assert self.binded, "call bind before initializing the parameters"
        if self.params_initialized and not force_init:
            return
        initializer = initializer or init_mod.Uniform(0.01)
        if arg_params is None and aux_params is None and getattr(self, "_preloaded", None) is not None:
            arg_params, aux_params = self._preloaded
        ex0 = self._execs[0]
        self._arg_params = {n: nd.zeros(ex0.arg_dict[n].shape) for n in self._param_names}
        self._aux_params = {n: nd.zeros(ex0.aux_dict[n].shape) for n in self._aux_names}
        attrs = self._symbol.attr_dict() if hasattr(self._symbol, "attr_dict") else {}
        for name, arr in list(self._arg_params.items()) + list(self._aux_params.items()):
            given = (arg_params or {}).get(name) if name in self._arg_params else (aux_params or {}).get(name)
            if given is not None:
                arr[:] = given
            elif (arg_params is not None or aux_params is not None) and not allow_missing and name in self._arg_params and arg_params is not None:
                raise MXNetError("%s is not presented" % name)
            else:
                initializer(init_mod.InitDesc(name, attrs.get(name)), arr)
        self.params_initialized = True
        self._sync_params_to_devices()

    def _sync_params_to_devices(self):
        for ex in self._execs:
            ex.copy_params_from(self._arg_params, self._aux_params, allow_extra_params=True)

    def get_params(self):
        assert self.binded and self.params_initialized
        ex0 = self._execs[0]                      # replicas are identical after update()
        for n in self._param_names:
            self._arg_params[n][:] = ex0.arg_dict[n].as_in_context(cpu())
        for n in self._aux_names:
            self._aux_params[n][:] = ex0.aux_dict[n].as_in_context(cpu())
        return self._arg_params, self._aux_params

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        self.init_params(None, arg_params, aux_params, allow_missing, force_init, allow_extra)

    # ---- optimizer + kvstore (module.py:472-560)
    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), force_init=False):
        assert self.binded and self.params_initialized
        if self.optimizer_initialized and not force_init:
            return
        kv, update_on_kvstore = _create_kvstore(kvstore, len(self._context), self._arg_params)
        batch_size = self._data_shapes[0].shape[0]
        if kv is not None and "dist" in kv.type and "_sync" in kv.type:
            batch_size *= kv.num_workers
        if isinstance(optimizer, str):
            params = dict(optimizer_params)
            params.setdefault("rescale_grad", 1.0 / batch_size)
            idx2name = {i: n for i, n in enumerate(self._param_names)} if update_on_kvstore else \
                {i * len(self._context) + k: n for i, n in enumerate(self._param_names) for k in range(len(self._context))}
            optimizer = opt.create(optimizer, param_idx2name=idx2name, **params)
        self._optimizer, self._kvstore, self._update_on_kvstore, self._updater = optimizer, kv, update_on_kvstore, None
        if kv is not None:
            _initialize_kvstore(kv, self._param_arrays(), self._arg_params, self._param_names, update_on_kvstore)
            if update_on_kvstore:
                kv.set_optimizer(optimizer)
        if not update_on_kvstore:
            self._updater = opt.get_updater(optimizer)
        self.optimizer_initialized = True

    def borrow_optimizer(self, shared_module):
        """Use the optimizer / updater / kvstore of ``shared_module`` (whose parameters this module shares)."""
        assert shared_module.optimizer_initialized
        self._optimizer, self._kvstore, self._update_on_kvstore, self._updater = (shared_module._optimizer, shared_module._kvstore,
                                                                                  shared_module._update_on_kvstore, shared_module._updater)
        self.optimizer_initialized = True

    def _param_arrays(self):
        return [[ex.arg_dict[n] for ex in self._execs] for n in self._param_names]

    def _grad_arrays(self):
        return [[ex.grad_dict.get(n) for ex in self._execs] for n in self._param_names]

    label_names = property(lambda self: self._label_names)
    data_shapes = property(lambda self: self._data_shapes)
    label_shapes = property(lambda self: self._label_shapes)

    @property
    def output_shapes(self):
        shapes = {d.name: d.shape for d in self._data_shapes + self._label_shapes}
        _, outs, _ = self._symbol.infer_shape(**shapes)
        return list(zip(self._symbol.list_outputs(), outs))

    def install_monitor(self, mon):
        """Attach a ``mx.monitor.Monitor`` to every executor."""
        assert self.binded
        for ex in self._execs:
            mon.install(ex)

    def reshape(self, data_shapes, label_shapes=None):
        """Re-bind for new input shapes; parameters keep their arrays (module.py:430-470)."""
        assert self.binded
        self._data_shapes, self._label_shapes = _as_desc(data_shapes), _as_desc(label_shapes)
        batch = self._data_shapes[0].shape[0]
        n = len(self._context)
        bounds = [(batch * i) // n for i in range(n + 1)]
        self._slices = [slice(bounds[i], bounds[i + 1]) for i in range(n)]
        self._execs = [ex.reshape(**{d.name: (sl.stop - sl.start,) + tuple(d.shape[1:]) for d in self._data_shapes + self._label_shapes})
                       for ex, sl in zip(self._execs, self._slices)]

    # ---- computation
    def forward(self, data_batch, is_train=None):
        assert self.binded and self.params_initialized
        new_shape = tuple(data_batch.data[0].shape) if hasattr(data_batch, "data") and data_batch.data else None
        if new_shape is not None and self._data_shapes and new_shape != tuple(self._data_shapes[0].shape):
            # a batch of a different shape re-binds on the fly (module.py:590-620)
            lab = getattr(data_batch, "label", None)
            self.reshape([(d.name, tuple(a.shape)) for d, a in zip(self._data_shapes, data_batch.data)],
                         [(d.name, tuple(a.shape)) for d, a in zip(self._label_shapes, lab)] if lab and self._label_shapes else None)
        is_train = self.for_training if is_train is None else is_train
        data = data_batch.data if isinstance(data_batch, DataBatch) or hasattr(data_batch, "data") else data_batch
        label = getattr(data_batch, "label", None)
        for ex, sl in zip(self._execs, self._slices):
            for name, arr in zip(self._data_names, data):
                ex.arg_dict[name][:] = arr[sl].as_in_context(ex._ctx)
            if label is not None:
                for name, arr in zip(self._label_names, label):
                    if name in ex.arg_dict:
                        ex.arg_dict[name][:] = arr[sl].as_in_context(ex._ctx)
            ex.forward(is_train=is_train)

    def backward(self, out_grads=None):
        assert self.binded and self.params_initialized and self.for_training
        for ex, sl in zip(self._execs, self._slices):
            g = None if out_grads is None else [o[sl] for o in (out_grads if isinstance(out_grads, (list, tuple)) else [out_grads])]
            ex.backward(g)

    def update(self):
        assert self.optimizer_initialized
        if self._update_on_kvstore:
            _update_params_on_kvstore(self._param_arrays(), self._grad_arrays(), self._kvstore, self._param_names)
        else:
            _update_params(self._param_arrays(), self._grad_arrays(), self._updater, len(self._context), self._kvstore, self._param_names)

    def get_outputs(self, merge_multi_context=True):
        outs = [ex.outputs for ex in self._execs]
        if not merge_multi_context:
            return outs
        if len(outs) == 1:
            return outs[0]
        return [nd.concat(*[o[i].as_in_context(outs[0][i].context) for o in outs], dim=0) for i in range(len(outs[0]))]

    def get_input_grads(self, merge_multi_context=True):
        assert self._inputs_need_grad
        gs = [[ex.grad_dict[n] for n in self._data_names] for ex in self._execs]
        if not merge_multi_context or len(gs) == 1:
            return gs[0] if len(gs) == 1 else gs
        return [nd.concat(*[g[i].as_in_context(gs[0][i].context) for g in gs], dim=0) for i in range(len(gs[0]))]

    def update_metric(self, eval_metric, labels):
        eval_metric.update(labels, self.get_outputs())

    # ---- checkpoints (module.py:161-203)
    def save_checkpoint(self, prefix, epoch, save_optimizer_states=False):
        arg, aux = self.get_params()
        save_checkpoint(prefix, epoch, self._symbol.tojson(), arg, aux)
        if save_optimizer_states:
            self.save_optimizer_states("%s-%04d.states" % (prefix, epoch))

    def save_optimizer_states(self, fname):
        assert self.optimizer_initialized
        if self._update_on_kvstore:
            self._kvstore.save_optimizer_states(fname)
        else:
            with open(fname, "wb") as f:
                f.write(self._updater.get_states())

    def load_optimizer_states(self, fname):
        assert self.optimizer_initialized
        if self._update_on_kvstore:
            self._kvstore.load_optimizer_states(fname)
        else:
            with open(fname, "rb") as f:
                self._updater.set_states(f.read())

    @staticmethod
    def load(prefix, epoch, load_optimizer_states=False, **kwargs):
        from . import symbol as sym
        net, arg, aux = load_checkpoint(prefix, epoch)
        mod = Module(net, **kwargs)
        mod._preloaded = (arg, aux)                # used by init_params() when no explicit values are given
        if load_optimizer_states:
            mod._preload_opt_states = "%s-%04d.states" % (prefix, epoch)
        return mod


class BucketingModule(BaseModule):
    """One Module per bucket key (e.g. sequence length), all sharing parameters, gradients and the optimizer with the default bucket's
    module (parity: python/mxnet/module/bucketing_module.py).  ``sym_gen(key) -> (symbol, data_names, label_names)``; a batch selects its
    bucket through ``batch.bucket_key`` (+ ``provide_data`` / ``provide_label`` for the shapes)."""

    def __init__(self, sym_gen, default_bucket_key=None, logger=logging, context=None, fixed_param_names=None):
        super().__init__(logger)
        assert default_bucket_key is not None
        self._sym_gen, self._default_key, self._context, self._fixed = sym_gen, default_bucket_key, context, fixed_param_names
        self._buckets, self._curr, self._curr_key = {}, None, None
        sym, dn, ln = sym_gen(default_bucket_key)
        self._data_names_, self._label_names_ = list(dn or []), list(ln or [])

    data_names = property(lambda self: self._data_names_)
    symbol = property(lambda self: self._curr.symbol)
    output_names = property(lambda self: self._curr.output_names)
    data_shapes = property(lambda self: self._curr.data_shapes)          # of the bucket that is currently switched in
    label_shapes = property(lambda self: self._curr.label_shapes)
    output_shapes = property(lambda self: self._curr.output_shapes)

    def _make(self, key):
        sym, dn, ln = self._sym_gen(key)
        return Module(sym, dn, ln, self.logger, self._context, self._fixed)

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False, shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        self.for_training, self._inputs_need_grad, self._grad_req = for_training, inputs_need_grad, grad_req
        mod = self._make(self._default_key)
        mod.bind(data_shapes, label_shapes, for_training, inputs_need_grad, force_rebind=True, grad_req=grad_req)
        self._buckets = {self._default_key: mod}
        self._curr, self._curr_key, self.binded = mod, self._default_key, True

    def switch_bucket(self, bucket_key, data_shapes, label_shapes=None):
        assert self.binded, "call bind before switching bucket"
        if bucket_key not in self._buckets:
            mod = self._make(bucket_key)
            mod.bind(data_shapes, label_shapes, self.for_training, self._inputs_need_grad, force_rebind=True,
                     shared_module=self._buckets[self._default_key], grad_req=self._grad_req)
            if self.optimizer_initialized:
                mod.borrow_optimizer(self._buckets[self._default_key])
            self._buckets[bucket_key] = mod
        self._curr, self._curr_key = self._buckets[bucket_key], bucket_key

    def init_params(self, *args, **kwargs):
        self._buckets[self._default_key].init_params(*args, **kwargs)
        self.params_initialized = True

    def get_params(self):
        return self._buckets[self._default_key].get_params()

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        self._buckets[self._default_key].set_params(arg_params, aux_params, allow_missing, force_init, allow_extra)
        self.params_initialized = True

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), force_init=False):
        if self.optimizer_initialized and not force_init:
            return
        base = self._buckets[self._default_key]
        base.init_optimizer(kvstore, optimizer, optimizer_params, force_init)
        for k, m in self._buckets.items():
            if m is not base:
                m.borrow_optimizer(base)
        self.optimizer_initialized = True

    def prepare(self, data_batch, sparse_row_id_fn=None):
        key = getattr(data_batch, "bucket_key", None)
        if key is not None:
            self.switch_bucket(key, data_batch.provide_data, data_batch.provide_label)

    def forward(self, data_batch, is_train=None):
        self.prepare(data_batch)
        self._curr.forward(data_batch, is_train)

    def backward(self, out_grads=None):
        self._curr.backward(out_grads)

    def update(self):
        self._curr.update()

    def get_outputs(self, merge_multi_context=True):
        return self._curr.get_outputs(merge_multi_context)

    def get_input_grads(self, merge_multi_context=True):
        return self._curr.get_input_grads(merge_multi_context)

    def update_metric(self, eval_metric, labels):
        self._curr.update_metric(eval_metric, labels)

    def save_checkpoint(self, prefix, epoch, save_optimizer_states=False):
        self._buckets[self._default_key].save_checkpoint(prefix, epoch, save_optimizer_states)


class SequentialModule(BaseModule):
    """A chain of modules: the outputs of module i are the data of module i+1, gradients flow back through ``get_input_grads``
    (parity: python/mxnet/module/sequential_module.py).  ``add(module, take_labels=True, auto_wiring=True)``."""
    META_TAKE_LABELS, META_AUTO_WIRING = "take_labels", "auto_wiring"

    def __init__(self, logger=logging):
        super().__init__(logger)
        self._modules, self._metas = [], []

    def add(self, module, **kwargs):
        self._modules.append(module); self._metas.append(kwargs)
        self.binded = self.params_initialized = self.optimizer_initialized = False
        return self

    data_names = property(lambda self: self._modules[0].data_names if self._modules else [])
    output_names = property(lambda self: self._modules[-1].output_names if self._modules else [])

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False, shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        assert self._modules, "Attempting to bind an empty SequentialModule"
        self.for_training, self._inputs_need_grad = for_training, inputs_need_grad
        self._label_shapes = label_shapes
        shapes = _as_desc(data_shapes)
        for i, (m, meta) in enumerate(zip(self._modules, self._metas)):
            lab = label_shapes if meta.get(self.META_TAKE_LABELS) else None
            m.bind(shapes, lab, for_training, inputs_need_grad or i > 0, force_rebind=True, grad_req=grad_req)
            if i + 1 < len(self._modules):
                nxt = self._modules[i + 1]
                _, out_shapes, _ = m.symbol.infer_shape(**{d.name: d.shape for d in _as_desc(shapes) + (_as_desc(lab) if lab else [])})
                names = nxt.data_names if self._metas[i + 1].get(self.META_AUTO_WIRING, True) else m.output_names
                shapes = [(n, s) for n, s in zip(names, out_shapes)]
        self.binded = True

    def init_params(self, initializer=None, arg_params=None, aux_params=None, allow_missing=False, force_init=False, allow_extra=False):
        for m in self._modules:
            m.init_params(initializer, arg_params, aux_params, allow_missing=True, force_init=force_init, allow_extra=True)
        self.params_initialized = True

    def get_params(self):
        arg, aux = {}, {}
        for m in self._modules:
            a, x = m.get_params(); arg.update(a); aux.update(x)
        return arg, aux

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), force_init=False):
        for m in self._modules:
            m.init_optimizer(kvstore, optimizer, optimizer_params, force_init)
        self.optimizer_initialized = True

    def forward(self, data_batch, is_train=None):
        batch = data_batch
        for i, (m, meta) in enumerate(zip(self._modules, self._metas)):
            m.forward(batch, is_train)
            if i + 1 < len(self._modules):
                batch = DataBatch(m.get_outputs(), data_batch.label if self._metas[i + 1].get(self.META_TAKE_LABELS) else None)

    def backward(self, out_grads=None):
        for i in range(len(self._modules) - 1, -1, -1):
            self._modules[i].backward(out_grads)
            if i > 0:
                out_grads = self._modules[i].get_input_grads()

    def update(self):
        for m in self._modules:
            m.update()

    def get_outputs(self, merge_multi_context=True):
        return self._modules[-1].get_outputs(merge_multi_context)

    def get_input_grads(self, merge_multi_context=True):
        return self._modules[0].get_input_grads(merge_multi_context)

    def update_metric(self, eval_metric, labels):
        for m, meta in zip(self._modules, self._metas):
            if meta.get(self.META_TAKE_LABELS):
                m.update_metric(eval_metric, labels)


class PythonModule(BaseModule):
    """Module whose computation is written in Python — no parameters, no optimizer (parity: python/mxnet/module/python_module.py).  Subclasses
    implement ``forward`` / ``backward`` / ``get_outputs`` / ``get_input_grads`` and ``_compute_output_shapes``."""

    def __init__(self, data_names, label_names, output_names, logger=logging):
        super().__init__(logger)
        self._data_names_, self._label_names_, self._output_names_ = list(data_names), list(label_names or []), list(output_names)
        self._data_shapes = self._label_shapes = self._output_shapes = None

    data_names = property(lambda self: self._data_names_)
    output_names = property(lambda self: self._output_names_)
    data_shapes = property(lambda self: self._data_shapes)
    label_shapes = property(lambda self: self._label_shapes)
    output_shapes = property(lambda self: self._output_shapes)

    def get_params(self):
        return {}, {}

    def init_params(self, *args, **kwargs):
        self.params_initialized = True

    def set_params(self, *args, **kwargs):
        self.params_initialized = True

    def update(self):
        pass

    def update_metric(self, eval_metric, labels):
        if self._label_shapes is not None:
            eval_metric.update(labels, self.get_outputs())

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False, shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        self.for_training, self._inputs_need_grad = for_training, inputs_need_grad
        self._data_shapes, self._label_shapes = _as_desc(data_shapes), (_as_desc(label_shapes) if label_shapes else None)
        self._output_shapes = self._compute_output_shapes()
        self.binded = True

    def _compute_output_shapes(self):
        raise NotImplementedError

    def init_optimizer(self, *args, **kwargs):
        self.optimizer_initialized = True


class PythonLossModule(PythonModule):
    """A loss layer in Python: forward stores the scores, backward produces ``grad_func(scores, labels)``; used as the last stage of a
    ``SequentialModule`` (python_module.py:240-340)."""

    def __init__(self, name="pyloss", data_names=("data",), label_names=("softmax_label",), logger=logging, grad_func=None):
        super().__init__(data_names, label_names, [name + "_output"], logger)
        self._name, self._scores, self._labels, self._scores_grad, self._grad_func = name, None, None, None, grad_func

    def _compute_output_shapes(self):
        return [(self._name + "_output", self._data_shapes[0].shape)]

    def forward(self, data_batch, is_train=None):
        self._scores = data_batch.data[0]
        if (self.for_training if is_train is None else is_train) and data_batch.label:
            self._labels = data_batch.label[0]

    def get_outputs(self, merge_multi_context=True):
        return [self._scores]

    def backward(self, out_grads=None):
        assert out_grads is None, "For a loss module, out_grads should be None"
        assert self.for_training
        if self._grad_func is None:
            raise NotImplementedError("pass grad_func or override backward")
        g = self._grad_func(self._scores, self._labels)
        self._scores_grad = g if isinstance(g, nd.NDArray) else nd.array(g)

    def get_input_grads(self, merge_multi_context=True):
        return [self._scores_grad]

    def install_monitor(self, mon):
        raise NotImplementedError


class DataParallelExecutorGroup:
    """The per-context executors of one symbol with batch slicing (``python/mxnet/module/executor_group.py``), for code that drives them
    without a Module.  Here the Module class itself owns exactly this machinery, so the group is a view onto a bound Module with the
    reference's constructor and method names."""

    def __init__(self, symbol, contexts, workload, data_shapes, label_shapes, param_names, for_training, inputs_need_grad, shared_group=None,
                 logger=logging, fixed_param_names=None, grad_req="write", state_names=None, group2ctxs=None):
        name_of = lambda d: d[0] if isinstance(d, (tuple, list)) else d.name
        self.symbol, self.contexts, self.workload = symbol, list(contexts), workload
        self.param_names, self.for_training, self.inputs_need_grad = list(param_names), for_training, inputs_need_grad
        self._mod = Module(symbol, data_names=[name_of(d) for d in data_shapes], label_names=[name_of(d) for d in (label_shapes or [])] or None,
                           logger=logger, context=self.contexts, fixed_param_names=fixed_param_names)
        self._mod.bind(data_shapes, label_shapes, for_training=for_training, inputs_need_grad=inputs_need_grad,
                       shared_module=shared_group._mod if shared_group is not None else None, grad_req=grad_req)
        self._mod.params_initialized = True          # parameters are whatever set_params puts there (zeros until then)

    execs = property(lambda self: self._mod._execs)
    slices = property(lambda self: self._mod._slices)
    data_shapes = property(lambda self: self._mod.data_shapes)
    label_shapes = property(lambda self: self._mod.label_shapes)
    param_arrays = property(lambda self: self._mod._param_arrays())
    grad_arrays = property(lambda self: self._mod._grad_arrays())
    aux_arrays = property(lambda self: [[ex.aux_dict[n] for ex in self._mod._execs] for n in self._mod._aux_names])

    def reshape(self, data_shapes, label_shapes):
        self._mod.reshape(data_shapes, label_shapes)

    def set_params(self, arg_params, aux_params, allow_extra=False):
        for ex in self._mod._execs:
            ex.copy_params_from(arg_params, aux_params, allow_extra_params=allow_extra)

    def get_params(self, arg_params, aux_params):
        """Write the current parameters (first device: replicas are identical) into the given dicts."""
        ex0 = self._mod._execs[0]
        for n in self.param_names:
            arg_params[n] = ex0.arg_dict[n].as_in_context(cpu()) if n not in arg_params else arg_params[n]
            arg_params[n][:] = ex0.arg_dict[n].as_in_context(cpu())
        for n in self._mod._aux_names:
            aux_params[n] = ex0.aux_dict[n].as_in_context(cpu()) if n not in aux_params else aux_params[n]
            aux_params[n][:] = ex0.aux_dict[n].as_in_context(cpu())

    def forward(self, data_batch, is_train=None):
        self._mod.forward(data_batch, is_train=is_train)

    def backward(self, out_grads=None):
        self._mod.backward(out_grads)

    def get_outputs(self, merge_multi_context=True, begin=0, end=None):
        outs = self._mod.get_outputs(merge_multi_context)
        return outs[begin:end]

    def get_input_grads(self, merge_multi_context=True):
        return self._mod.get_input_grads(merge_multi_context)

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        self._mod.update_metric(eval_metric, labels)

    def install_monitor(self, mon):
        self._mod.install_monitor(mon)


# the reference's package layout (python/mxnet/module/*.py) as importable paths
def _register_paths():
    from ._alias import submodule
    submodule(__name__, "base_module", {"BaseModule": BaseModule})
    submodule(__name__, "module", {"Module": Module})
    submodule(__name__, "bucketing_module", {"BucketingModule": BucketingModule})
    submodule(__name__, "sequential_module", {"SequentialModule": SequentialModule})
    submodule(__name__, "python_module", {"PythonModule": PythonModule, "PythonLossModule": PythonLossModule})
    submodule(__name__, "executor_group", {"DataParallelExecutorGroup": DataParallelExecutorGroup})


_register_paths()
