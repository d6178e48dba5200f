from . import transforms  # noqa: F401
from .datasets import *  # noqa: F401,F403
