import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "support")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"
HAVE_REFERENCE = os.path.isdir(REFERENCE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference .tla files are not on this machine")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def registry():
    with open(os.path.join(ROOT, "models", "MODELS.json")) as f:
        return json.load(f)
