import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_DATA = "/root/reference/tests/data"  # only exists in the build container


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gb():
    """galah_b200 bound to cuda:0 (GPU tests only). Fails loudly if the .so or GPU is missing."""
    import galah_b200
    galah_b200.init(0)
    return galah_b200


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "fixture_sketches.npz"))
