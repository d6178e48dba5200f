"""SVRGModule (Johnson & Zhang 2013).  Every ``update_freq`` epochs a snapshot ``w~`` of the weights is taken and the FULL gradient
``mu = 1/n sum_i grad f_i(w~)`` is computed over the training set; each step then uses ``g = grad f_i(w) - grad f_i(w~) + mu``.

Parity: ``svrg_module.py:30-560`` — same public surface (``update_full_grads``, ``fit`` with the periodic refresh, ``forward``/``backward``/
``update`` driving a second executor group bound to the snapshot weights).  The reference routes the three gradient terms through
dedicated kvstore keys and an ``_SVRGOptimizer`` wrapper that re-assigns them; here the combination is done in place on the executors'
gradient arrays right before the regular ``Module.update`` (which then goes through whatever kvstore / optimizer is configured, HiPS
included), so no wrapper optimizer is needed."""
from __future__ import annotations

import logging
import time

from ... import metric as _metric
from ... import ndarray as nd
from ...module import Module

__all__ = ["SVRGModule"]


class SVRGModule(Module):
    def __init__(self, symbol, data_names=("data",), label_names=("softmax_label",), logger=logging, context=None, fixed_param_names=None,
                 update_freq=None, **kwargs):
        super().__init__(symbol, data_names=data_names, label_names=label_names, logger=logger, context=context, fixed_param_names=fixed_param_names)
        if not isinstance(update_freq, int) or update_freq <= 0:
            raise ValueError("update_freq in SVRGModule must be a positive integer to represent the frequency for calculating full gradients")
        self.update_freq = update_freq
        self._mod_aux = Module(symbol, data_names, label_names, logger, context, fixed_param_names)
        self._full_grads = None                                        # {param name: NDArray}, averaged over the whole data set

    # ---- the snapshot module mirrors every structural call
    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False, shared_module=None, grad_req="write"):
        super().bind(data_shapes, label_shapes, for_training, inputs_need_grad, force_rebind, shared_module, grad_req)
        if for_training:
            self._mod_aux.bind(data_shapes, label_shapes, for_training, inputs_need_grad, force_rebind, shared_module, grad_req)

    def reshape(self, data_shapes, label_shapes=None):
        self.bind(data_shapes, label_shapes, self.for_training, self._inputs_need_grad, force_rebind=True)

    def init_params(self, *args, **kwargs):
        super().init_params(*args, **kwargs)
        if self._mod_aux.binded:
            arg, aux = self.get_params()
            self._mod_aux.init_params(arg_params=arg, aux_params=aux, allow_missing=False, force_init=True)

    def forward(self, data_batch, is_train=None):
        super().forward(data_batch, is_train)
        if (self.for_training if is_train is None else is_train) and self._mod_aux.binded:
            self._mod_aux.forward(data_batch, is_train=True)

    def backward(self, out_grads=None):
        super().backward(out_grads)
        if self._mod_aux.binded:
            self._mod_aux.backward(out_grads)

    def update(self):
        self._svrg_grads_update_rule()
        super().update()

    def _svrg_grads_update_rule(self):
        """``g <- g(w) - g(w~) + mu`` on every executor's gradient arrays."""
        if self._full_grads is None:
            return
        for n in self._param_names:
            for ex, ex_aux in zip(self._execs, self._mod_aux._execs):
                g = ex.grad_dict.get(n)
                if g is None:
                    continue
                g._t.sub_(ex_aux.grad_dict[n]._t).add_(self._full_grads[n]._t.to(g._t.device) / len(self._execs))

    def update_full_grads(self, train_data):
        """Snapshot the weights into the auxiliary module and average its gradients over ALL batches of ``train_data``."""
        arg, aux = self.get_params()
        self._mod_aux.set_params(arg_params=arg, aux_params=aux)
        train_data.reset()
        acc, nbatch = {n: None for n in self._param_names}, 0
        for batch in train_data:
            self._mod_aux.forward(batch, is_train=True)
            self._mod_aux.backward()
            nbatch += 1
            for n in self._param_names:
                gs = [ex.grad_dict[n] for ex in self._mod_aux._execs if ex.grad_dict.get(n) is not None]
                if not gs:
                    continue
                tot = gs[0]._t.clone()
                for g in gs[1:]:
                    tot += g._t.to(tot.device)
                acc[n] = tot if acc[n] is None else acc[n] + tot
        self._full_grads = {n: nd.NDArray(t / max(nbatch, 1)) for n, t in acc.items() if t is not None}
        train_data.reset()

    def fit(self, train_data, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None, kvstore="local",
            optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), eval_end_callback=None, eval_batch_end_callback=None, initializer=None,
            arg_params=None, aux_params=None, allow_missing=False, force_rebind=False, force_init=False, begin_epoch=0, num_epoch=None,
            validation_metric=None, monitor=None, sparse_row_id_fn=None):
        assert num_epoch is not None, "please specify number of epochs"
        self.bind(train_data.provide_data, train_data.provide_label, for_training=True, force_rebind=force_rebind)
        self.init_params(initializer=initializer, arg_params=arg_params, aux_params=aux_params, allow_missing=allow_missing, force_init=force_init)
        self.init_optimizer(kvstore=kvstore, optimizer=optimizer, optimizer_params=optimizer_params)
        if not isinstance(eval_metric, _metric.EvalMetric):
            eval_metric = _metric.create(eval_metric)
        validation_metric = validation_metric or eval_metric
        from ...model import BatchEndParam
        for epoch in range(begin_epoch, num_epoch):
            tic = time.time()
            eval_metric.reset()
            if epoch % self.update_freq == 0:
                self.update_full_grads(train_data)
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
                res = self.score(eval_data, validation_metric, epoch=epoch)
                for name, val in res:
                    self.logger.info("Epoch[%d] Validation-%s=%f", epoch, name, val)
