import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA GPU (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu(n): test needs n GPUs")
    config.addinivalue_line("markers", "world_size(n): reference-style marker (informational)")


def pytest_collection_modifyitems(config, items):
    import torch
    have = torch.cuda.device_count() if torch.cuda.is_available() else 0
    for item in items:
        if "gpu" in item.keywords and have == 0:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        m = item.get_closest_marker("multigpu")
        if m is not None and have < int(m.args[0]):
            item.add_marker(pytest.mark.skip(reason=f"needs {m.args[0]} GPUs, have {have}"))
