import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """compiled reference oracle (oracle/_ref); parity tests skip loudly if it did not travel"""
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libknowhere_ref.so not built")
    return r


@pytest.fixture(scope="session")
def kb():
    import knowhere_b200
    return knowhere_b200
