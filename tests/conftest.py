import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_share():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "share_small.json")))


@pytest.fixture(scope="session")
def verifier():
    import dvt_circuits_b200 as dk
    v = dk.Verifier(0)
    yield v
    v.close()
