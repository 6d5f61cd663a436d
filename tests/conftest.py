import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(autouse=True)
def _seed_global_rngs():
    """Variables initialise from numpy's global RNG (tf.truncated_normal in the reference, param.py:206-208) and the
    Indexer draws minibatches from it (model.py:147-149): seed it per test so that every test sees the same model whatever
    the order / selection of tests in the process (unseeded, borderline tolerances on the ill-conditioned notebook
    inputs flaked from run to run)."""
    import numpy as np
    np.random.seed(1234)
    try:
        import torch
        torch.manual_seed(1234)
    except Exception:  # pragma: no cover
        pass
    yield


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must fail loudly on a GPU box if the extension is missing; on a CPU-only box they
    # are deselected by `-m "not gpu"`.  If someone runs them without a GPU, skip with a clear reason.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
