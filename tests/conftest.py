import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


SMALL_BINS = [60, 55, 50, 48, 45, 43, 40, 37, 36, 34, 34, 33, 29, 27, 26, 23, 20, 20, 15, 16, 12, 13]


@pytest.fixture(scope="session")
def small_bins():
    return list(SMALL_BINS)
