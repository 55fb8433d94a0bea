import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def meshes():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes.npz"))
    return {k[:-2]: (z[k[:-2] + "_v"], z[k[:-2] + "_f"]) for k in z.files if k.endswith("_v")}
