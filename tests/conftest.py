import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_frontend():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "frontend.json")) as f:
        return json.load(f)["entries"]


@pytest.fixture(scope="session")
def fixtures():
    """The reference's bundled 4-row fixtures (data/test.csv, data/extended.csv), restated as arrays."""
    import numpy as np
    return {
        "test": {"price": np.array([10.5, 20.0, 15.25, 30.0], np.float32),
                 "quantity": np.array([3, 4, 2, 5], np.int32)},
        "extended": {"price": np.array([10.5, 20.0, 15.25, 30.0], np.float32),
                     "quantity": np.array([3, 4, 2, 5], np.int32),
                     "discount": np.array([0.1, 0.2, 0.05, 0.15], np.float32)},
    }
