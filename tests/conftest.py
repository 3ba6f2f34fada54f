import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden_v1.json")) as fh:
        meta = json.load(fh)
    arrays = dict(np.load(os.path.join(GOLDEN_DIR, "golden_v1.npz")))
    return meta, arrays


@pytest.fixture(scope="session")
def qw():
    """The product package (CUDA path)."""
    import qwtdh_b200
    return qwtdh_b200
