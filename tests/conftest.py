import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_maps():
    out = {}
    out.update(load_golden("maps_v0_1000_1099.json.gz"))
    out.update(load_golden("maps_misc.json.gz"))
    return out


@pytest.fixture(scope="session")
def golden_resets():
    out = {}
    out.update(load_golden("reset_v0_1000_1099.json.gz"))
    out.update(load_golden("reset_misc.json.gz"))
    return out
