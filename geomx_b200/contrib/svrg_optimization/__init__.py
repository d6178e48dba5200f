"""``mx.contrib.svrg_optimization`` — stochastic variance-reduced gradient training for Modules
(parity: python/mxnet/contrib/svrg_optimization/{svrg_module,svrg_optimizer}.py)."""
from .svrg_module import SVRGModule  # noqa: F401
