"""Base definitions shared by the whole package.

Parity: ``python/mxnet/base.py`` (MXNetError, string/numeric type tuples) in the
reference.  Nothing here touches a GPU.
"""
from __future__ import annotations

import os

__all__ = ["MXNetError", "GeoMXError", "string_types", "numeric_types", "integer_types", "getenv_int",
           "getenv_float", "getenv_str", "getenv_bool"]


class MXNetError(RuntimeError):
    """Error raised by the framework (name kept for script compatibility)."""


GeoMXError = MXNetError

string_types = (str,)
integer_types = (int,)
numeric_types = (float, int)
try:  # numpy scalars behave as numbers everywhere in the API
    import numpy as _np

    integer_types = (int, _np.integer)
    numeric_types = (float, int, _np.generic)
except Exception:  # pragma: no cover
    pass


def getenv_str(name: str, default: str = "") -> str:
    v = os.environ.get(name)
    return default if v is None or v == "" else v


def getenv_int(name: str, default: int = 0) -> int:
    v = os.environ.get(name)
    if v is None or v == "":
        return default
    try:
        return int(float(v))
    except ValueError:
        return default


def getenv_float(name: str, default: float = 0.0) -> float:
    v = os.environ.get(name)
    if v is None or v == "":
        return default
    try:
        return float(v)
    except ValueError:
        return default


def getenv_bool(name: str, default: bool = False) -> bool:
    v = os.environ.get(name)
    if v is None or v == "":
        return default
    return v.strip().lower() not in ("0", "false", "no", "off")
