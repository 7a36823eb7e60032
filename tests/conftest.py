import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    import oracle

    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle

    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libhimg_ref.so not available")
    return oracle.ref()


@pytest.fixture(scope="session")
def fixtures():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))


@pytest.fixture(scope="session")
def golden_hashes():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "golden_hashes.json")) as f:
        return json.load(f)
