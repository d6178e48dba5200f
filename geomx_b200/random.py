"""``mx.random`` — global seeding and the sampling shortcuts (parity: python/mxnet/random.py: ``seed(seed_state, ctx)`` seeds every device
generator; ``uniform`` / ``normal`` / ``randint`` / ``shuffle`` forward to ``mx.nd.random``)."""
from .ndarray.random import normal, randint, randn, seed, shuffle, uniform  # noqa: F401

__all__ = ["seed", "uniform", "normal", "randn", "randint", "shuffle"]
