import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def encoder_golden():
    import numpy as np
    g = np.load(os.path.join(GOLDEN, "encoder_golden.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def cnn_golden():
    import numpy as np
    g = np.load(os.path.join(GOLDEN, "cnn_golden.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def synthetic_weights():
    from svision_b200 import weights
    return weights.synthetic_weights()
