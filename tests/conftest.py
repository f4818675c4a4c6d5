import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def two_string_npy():
    # byte-identical copy of the reference fixture test_data/two_string.npy (106 bytes);
    # tests/golden/make_golden.py records how it was produced and checked.
    return os.path.join(GOLDEN, "two_string.npy")
