import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def ctx():
    """A tbv_ctx on cuda:0 through the C-ABI. Fails loudly when the library or the GPU is missing."""
    from tbv_slam_public_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def stream8():
    from tbv_slam_public_b200 import synth
    return synth.make_stream(8)
