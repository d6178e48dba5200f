"""``mx.contrib.autograd`` — the pre-gluon autograd API (parity: python/mxnet/contrib/autograd.py:32-230): ``set_is_training``,
``train_section`` / ``test_section``, ``mark_variables``, ``backward``, ``compute_gradient``, ``grad_and_loss``, ``grad``."""
from __future__ import annotations

import functools

from .. import autograd as _ag
from ..ndarray import NDArray, zeros_like

__all__ = ["set_is_training", "TrainingStateScope", "train_section", "test_section", "mark_variables", "backward", "compute_gradient",
           "grad_and_loss", "grad"]


def set_is_training(is_train):
    """Sets BOTH recording and training mode (the legacy API had a single switch); returns the previous state."""
    prev = _ag.is_recording()
    _ag.set_recording(bool(is_train)); _ag.set_training(bool(is_train))
    return prev


class TrainingStateScope:
    def __init__(self, enter_state):
        self._enter, self._prev = enter_state, None

    def __enter__(self):
        self._prev = set_is_training(self._enter)

    def __exit__(self, *exc):
        if self._prev != self._enter:
            set_is_training(self._prev)


def train_section():
    return TrainingStateScope(True)


def test_section():
    return TrainingStateScope(False)


def mark_variables(variables, gradients, grad_reqs="write"):
    return _ag.mark_variables(variables, gradients, grad_reqs)


def backward(outputs, out_grads=None, retain_graph=False):
    return _ag.backward(outputs, out_grads, retain_graph)


def compute_gradient(outputs):
    return backward(outputs)


def grad_and_loss(func, argnum=None):
    """Decorated ``func(*args) -> loss``; the wrapper returns ``(list of gradients w.r.t. the chosen args, loss)``."""
    @functools.wraps(func)
    def wrapped(*args):
        variables = list(args)
        if argnum is not None:
            idx = [argnum] if isinstance(argnum, int) else list(argnum)
            variables = [args[i] for i in idx]
        for x in variables:
            assert isinstance(x, NDArray), "type of autograd input should be NDArray"
        grads = [zeros_like(x) for x in variables]
        mark_variables(variables, grads)
        with train_section():
            outputs = func(*args)
        compute_gradient([outputs] if isinstance(outputs, NDArray) else outputs)
        return [v.grad for v in variables], outputs
    return wrapped


def grad(func, argnum=None):
    gl = grad_and_loss(func, argnum)

    @functools.wraps(gl)
    def wrapped(*args):
        return gl(*args)[0]
    return wrapped
