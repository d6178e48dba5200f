"""Utilities: per-iteration timing ``Measure`` (parity ``examples/utils.py:120-192``), env/flag registry."""
from .measure import Measure  # noqa: F401
from . import flags  # noqa: F401
