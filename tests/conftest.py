import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    try:
        oracle.lib()
    except Exception as e:  # pragma: no cover
        pytest.skip(f"oracle library unavailable: {e}")
    return oracle
