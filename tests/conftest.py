import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "wekws"))


def reference_init_model():
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from wekws.model.kws_model import init_model
    return init_model


@pytest.fixture(scope="session")
def native():
    from wekws_b200 import _native
    return _native
