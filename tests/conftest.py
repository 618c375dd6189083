import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The tight-parity tests run the fp32 CUDA-core contractions; tests/test_gpu_tf32.py covers the tcgen05
# (kind::tf32) path, which is the library default, with its own stated tolerances.
os.environ.setdefault("BSED_PRECISION", "fp32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def has_reference():
    return os.path.isdir("/root/reference/src/models")
