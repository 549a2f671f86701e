import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", params=["init", "perturbed"])
def style(request):
    return request.param


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(style):
        if style not in cache:
            cache[style] = dict(np.load(os.path.join(GOLDEN_DIR, "v2_%s.npz" % style)))
        return cache[style]
    return load
