"""Parallelism: HiPS topology, symmetric-memory fabric (NVSwitch data plane), NCCL oracle."""
from .arena import ArenaLayout, TILE  # noqa: F401
from .fabric import HipsFabric, SymmetricHeap, Topology  # noqa: F401
