import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GEOMX_SYNTHETIC_SIZE", "2048")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 GPUs")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    ngpu = torch.cuda.device_count() if has_gpu else 0
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_mg = pytest.mark.skip(reason="needs >= 2 GPUs")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(skip_mg)
