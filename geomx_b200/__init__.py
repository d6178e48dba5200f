"""geomx_b200 — a Blackwell-native hierarchical-parameter-server training framework with GeoMX's capabilities.

``import geomx_b200 as mx`` gives the MXNet-shaped surface the reference's scripts use (``mx.nd``, ``mx.autograd``,
``mx.gluon``, ``mx.kv``, ``mx.optimizer``, ``mx.init``, ``mx.profiler`` …; reference ``python/mxnet/__init__.py``).
Like the reference (``python/mxnet/__init__.py:57`` → ``kvstore_server._init_kvstore_server_module``), importing the
package in a process whose ``DMLC_ROLE`` is *server* / *scheduler* (or global scheduler) turns that process into the
corresponding HiPS node and exits when the job finishes.
"""
from __future__ import annotations

__version__ = "0.1.0"

from . import base  # noqa: F401
from .base import MXNetError  # noqa: F401
from .context import Context, cpu, cpu_pinned, cpu_shared, current_context, gpu, num_gpus  # noqa: F401
from . import ndarray  # noqa: F401
from . import ndarray as nd  # noqa: F401
from . import autograd  # noqa: F401
from . import initializer  # noqa: F401
from . import initializer as init  # noqa: F401
from . import lr_scheduler  # noqa: F401
from . import optimizer  # noqa: F401
from . import ops  # noqa: F401
from . import kvstore  # noqa: F401
from . import kvstore as kv  # noqa: F401
from . import gluon  # noqa: F401
from . import metric  # noqa: F401
from . import model  # noqa: F401
from . import name  # noqa: F401
from . import attribute  # noqa: F401
from .attribute import AttrScope  # noqa: F401
from . import symbol  # noqa: F401
from . import symbol as sym  # noqa: F401
from . import module  # noqa: F401
from . import module as mod  # noqa: F401
from . import callback  # noqa: F401
from . import monitor  # noqa: F401
from . import test_utils  # noqa: F401
from . import operator  # noqa: F401
from . import random  # noqa: F401
from . import visualization  # noqa: F401
from . import visualization as viz  # noqa: F401
from . import image  # noqa: F401
from . import image as img  # noqa: F401
from . import rtc  # noqa: F401
from . import profiler  # noqa: F401
from . import io  # noqa: F401
from . import recordio  # noqa: F401
from . import storage  # noqa: F401
if base.getenv_str("GEOMX_GPU_MEM_POOL", "torch") == "native" and __import__("torch").cuda.is_available():       # before the first CUDA allocation
    storage.use_native_gpu_pool()
from . import engine  # noqa: F401
from . import utils  # noqa: F401
from . import parallel  # noqa: F401
from . import models  # noqa: F401
from . import contrib  # noqa: F401
from . import executor, executor_manager, libinfo, log, misc, notebook, registry, rnn, util  # noqa: F401
from . import kvstore_server  # noqa: F401
from . import monitor as mon  # noqa: F401
from . import random as rnd  # noqa: F401
from . import ndarray_doc, symbol_doc  # noqa: F401
from . import torch  # noqa: F401
from . import torch as th  # noqa: F401
from . import predictor  # noqa: F401

# server / scheduler bootstrap on import (no-op for workers and plain library use)
kvstore_server._init_kvstore_server_module()
