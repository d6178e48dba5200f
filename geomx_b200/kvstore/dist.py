"""Distributed KVStore front-end (filled in by the HiPS runtime modules)."""
from __future__ import annotations

import os

from ..base import MXNetError


def create_dist(name):
    fabric = os.environ.get("GEOMX_FABRIC", "auto").lower()
    has_ps_env = any(k in os.environ for k in ("DMLC_PS_ROOT_URI", "DMLC_ROLE", "DMLC_ROLE_GLOBAL", "DMLC_PS_GLOBAL_ROOT_URI"))
    if fabric in ("symm", "nccl") or (fabric == "auto" and not has_ps_env and "RANK" in os.environ):
        from ..parallel.fabric_kvstore import KVStoreFabric
        return KVStoreFabric(name)
    if fabric == "auto" and not has_ps_env:
        # single process, no launcher: degenerate 1-party / 1-worker HiPS on the in-process fabric
        from ..parallel.fabric_kvstore import KVStoreFabric
        return KVStoreFabric(name)
    from .dist_ps import KVStoreDist
    return KVStoreDist(name)
