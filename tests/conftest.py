import os
import sys

import pytest

# a dead peer in a collective test should fail the test in seconds, not after the production timeout
os.environ.setdefault("LDB_PEER_TIMEOUT_MS", "5000")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): reference-compiled runtime when oracle/_ref exists, else the port."""
    from oracle import oracle as O
    return O.Oracle("auto", workers=4)


@pytest.fixture(scope="session")
def gpu_ctx():
    from lingodb_b200 import runtime
    ctx = runtime.Context(0)
    yield ctx
    ctx.close()
