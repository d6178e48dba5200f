"""``mx.optimizer.contrib`` — contributed optimizers (reference: ``python/mxnet/optimizer/contrib.py``)."""
import torch

from ..base import MXNetError
from ..ndarray import NDArray
from .optimizer import Optimizer, register

__all__ = ["GroupAdaGrad"]


@register
class GroupAdaGrad(Optimizer):
    """AdaGrad with ONE accumulator per row of a 2-D parameter (embedding tables): ``history[r] += mean(grad[r]^2)``,
    ``weight[r] -= lr * grad[r] / sqrt(history[r] + eps)``.  A ``row_sparse`` gradient touches only its rows (lazy update, through
    ``mx.nd.contrib.group_adagrad_update``).  Weight decay is not supported, as in the reference."""

    def __init__(self, eps=1e-5, **kwargs):
        super().__init__(**kwargs)
        self.float_stable_eps = eps

    def create_state(self, index, weight):
        if len(weight.shape) != 2:
            raise MXNetError("GroupAdaGrad needs 2-D parameters, got shape %s" % (tuple(weight.shape),))
        t = weight._t if isinstance(weight, NDArray) else weight
        return NDArray(torch.zeros(t.shape[0], 1, dtype=t.dtype, device=t.device))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        if wd != 0:
            raise MXNetError("Weight decay is not supported for GroupAdaGrad")
        if getattr(grad, "stype", "default") == "row_sparse":
            from ..ndarray import contrib as ndc
            ndc.group_adagrad_update(weight, grad, state, lr=lr, rescale_grad=self.rescale_grad,
                                     clip_gradient=-1.0 if self.clip_gradient is None else self.clip_gradient, epsilon=self.float_stable_eps, out=weight)
            return
        w, h = weight._t, state._t
        g = grad._t * self.rescale_grad
        if self.clip_gradient is not None:
            g = g.clamp(-self.clip_gradient, self.clip_gradient)
        h.add_((g * g).mean(dim=1, keepdim=True))
        w.sub_(lr * g / torch.sqrt(h + self.float_stable_eps))
