"""``mx.gluon`` namespace."""
from . import data, loss, nn, utils  # noqa: F401
from .block import Block, HybridBlock, SymbolBlock  # noqa: F401
from .parameter import Constant, DeferredInitializationError, Parameter, ParameterDict  # noqa: F401
from .trainer import Trainer  # noqa: F401

from . import contrib  # noqa: F401,E402
from . import model_zoo  # noqa: F401,E402
from . import rnn  # noqa: F401,E402
