import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import build
    return build.build()


@pytest.fixture(scope="session")
def ctx_small():
    """A context sized for the parity-test images (<= 1920x1080)."""
    from polychase_b200 import capi
    c = capi.Context(max_width=1920, max_height=1088, max_features=16384)
    yield c
    c.close()
