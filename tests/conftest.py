import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def bfv_input():
    from oracle import bfv
    return bfv.load_input(os.path.join(GOLDEN, "bfv.in"))


@pytest.fixture(scope="session")
def bfv_empty_input():
    from oracle import bfv
    return bfv.load_input(os.path.join(GOLDEN, "bfv_empty.in"))


@pytest.fixture(scope="session")
def digests():
    return json.load(open(os.path.join(GOLDEN, "oracle_digests.json")))


@pytest.fixture(scope="session")
def golden_gamma(digests):
    return int(digests["gamma"], 16)


@pytest.fixture(scope="session")
def oracle_tables(bfv_input, golden_gamma):
    """Oracle advice tables for bfv.in under the golden gamma (≈5 s, shared)."""
    from oracle import bfv
    return bfv.build_tables(bfv_input, golden_gamma)
