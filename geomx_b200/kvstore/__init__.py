"""``mx.kv`` / ``mx.kvstore`` namespace."""
from . import compression  # noqa: F401
from .base import KVStoreBase, create, flush_all  # noqa: F401
from .base import KVStoreBase as KVStore  # noqa: F401
