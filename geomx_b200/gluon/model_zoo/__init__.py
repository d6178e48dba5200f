"""``mx.gluon.model_zoo`` (parity: python/mxnet/gluon/model_zoo): model definitions built from this framework's layers.  ``pretrained=True``
is not available offline; ``get_model(name, **kwargs)`` returns randomly initialisable networks."""
from . import vision  # noqa: F401
from .vision import get_model  # noqa: F401
from . import model_store  # noqa: F401,E402
