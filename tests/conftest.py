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


# The driver runs the GPU suite with -x.  Tests of the newest, least exercised features (spill, checkpoint/recover,
# -tool output, anything added after the round's last GPU pass) are collected LAST, so that a failure in one of
# them cannot hide the verdicts of the parity tests proper.
RUN_LAST = ("test_spill_store_smaller_than_the_state_space", "test_error_trace_through_spilled_levels",
            "test_checkpoint_and_recover", "test_cli_tool_mode_wraps_messages_in_tlc_markers", "test_zz_")


def pytest_collection_modifyitems(config, items):
    late = [it for it in items if it.name.startswith(RUN_LAST) or it.fspath.basename.startswith("test_zz_")]
    if late:
        ids = {id(it) for it in late}
        # among the late ones: first those whose expected results are fully established on the CPU (test_zz_*: models
        # that agree with the oracles state for state on the host), last the features that only a GPU can exercise
        # at all (spill ring, checkpoint / recover)
        late.sort(key=lambda it: 0 if it.fspath.basename.startswith("test_zz_") else 1)
        items[:] = [it for it in items if id(it) not in ids] + late


needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference .tla files are not on this machine")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def registry():
    with open(os.path.join(ROOT, "models", "MODELS.json")) as f:
        return json.load(f)
