"""Name-compatible entry point for scripts written against the reference's ``examples/utils.py`` (samplers, data loading, evaluation,
``Measure``): everything lives in ``examples/common.py`` and ``geomx_b200.utils.measure``."""
from common import *  # noqa: F401,F403
from common import _Slice as SplitSampler  # noqa: F401  (reference name: examples/utils.py SplitSampler / ClassSplitSampler)
from geomx_b200.utils.measure import Measure  # noqa: F401

eval_acc = accuracy  # noqa: F405  (reference name)
