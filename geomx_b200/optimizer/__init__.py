"""``mx.optimizer`` namespace."""
from .optimizer import *  # noqa: F401,F403
from .optimizer import Optimizer, Updater, create, get_updater, register  # noqa: F401
from . import contrib  # noqa: F401,E402
from .contrib import GroupAdaGrad  # noqa: F401,E402
