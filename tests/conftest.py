import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_SO = os.path.join(ROOT, "oracle", "libmeepo_oracle.so")
PRODUCT_SO = os.path.join(ROOT, "meepoembedding_b200", "libmeepo.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The authored CPU oracle (test infrastructure). Built on demand where a compiler exists."""
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    from meepoembedding_b200 import load_library

    return load_library(ORACLE_SO)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly (no skip, no fallback) if it is missing."""
    from meepoembedding_b200 import product_library

    return product_library()
