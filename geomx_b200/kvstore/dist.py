"""Distributed KVStore front-end (filled in by the HiPS runtime modules)."""
from __future__ import annotations

import os



def create_dist(name):
    fabric = os.environ.get("GEOMX_FABRIC", "auto").lower()
    has_ps_env = any(k in os.environ for k in ("DMLC_PS_ROOT_URI", "DMLC_ROLE", "DMLC_ROLE_GLOBAL", "DMLC_PS_GLOBAL_ROOT_URI"))
    if has_ps_env and os.environ.get("DMLC_ROLE") == "worker" and int(os.environ.get("WORLD_SIZE", "1")) > 1 and fabric == "auto":
        # torchrun inside a box + the reference's worker environment for the box: several boxes, one TCP endpoint per box (kvstore/hybrid.py)
        from .hybrid import KVStoreHybrid
        return KVStoreHybrid(name)
    if fabric == "gloo":
        from .collective import KVStoreCollective
        return KVStoreCollective(name)
    if fabric in ("symm", "nccl") or (fabric == "auto" and not has_ps_env and "RANK" in os.environ):
        import torch
        if fabric == "auto" and not torch.cuda.is_available():
            # torchrun on a machine without GPUs: same semantics over gloo collectives (kvstore/collective.py)
            from .collective import KVStoreCollective
            return KVStoreCollective(name)
        from ..parallel.fabric_kvstore import KVStoreFabric
        return KVStoreFabric(name)
    if fabric == "auto" and not has_ps_env:
        import torch
        if torch.cuda.is_available():
            # single process, no launcher: degenerate 1-party / 1-worker HiPS on the in-process fabric
            from ..parallel.fabric_kvstore import KVStoreFabric
            return KVStoreFabric(name)
        return KVStoreDistSingle(name)
    from .dist_ps import KVStoreDist
    return KVStoreDist(name)


from .local import KVStoreLocal  # noqa: E402


class KVStoreDistSingle(KVStoreLocal):
    """``dist_*`` in ONE CPU process without any launcher environment: both PS tiers collapse into the local store (1 worker, the optimizer
    runs in-process).  Lets the demo scripts run stand-alone on a machine without GPUs."""

    def __init__(self, name):
        super().__init__("local")
        self._type = name

    @property
    def configures_servers(self):
        return True

    def _set_gradient_compression(self, params):
        # single tier, single worker: there is no inter-tier link to compress; the setting is accepted and recorded
        self._gc.set_params(params)
