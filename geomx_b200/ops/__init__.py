"""Operators: PyTorch definitions (CPU / oracle) and hand-written sm_100a kernels (CUDA)."""
from . import functional, native  # noqa: F401
