import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs through the C ABI of libgisnav_b200.so)")


@pytest.fixture(scope="session")
def rand_params():
    from gisnav_b200 import weights as W

    return W.unpack(W.pack(W.random_init(0)))


@pytest.fixture(scope="session")
def rand_blob():
    from gisnav_b200 import weights as W

    return W.pack(W.random_init(0))


@pytest.fixture(scope="session")
def stages():
    return dict(np.load(os.path.join(GOLDEN, "stages_small.npz")))


def golden_pnp(seed):
    return dict(np.load(os.path.join(GOLDEN, f"pnp_cv2_{seed}.npz")))
