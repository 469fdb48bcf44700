import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Everything native is built in-tree before any test (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def search_golden():
    return load_golden("search_golden.json")


@pytest.fixture(scope="session")
def pattern_golden():
    return load_golden("pattern_golden.json")


@pytest.fixture(scope="session")
def request_golden():
    return load_golden("request_golden.json")


@pytest.fixture(scope="session")
def engine():
    import patmatchdocker_b200 as pm
    eng = pm.Engine(0)
    yield eng
    eng.close()
