import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Everything native is built in-tree (idempotent; on the GPU box the prebuilt files are used as they are)."""
    from flye_b200 import build
    if not os.path.exists(os.path.join(ROOT, "flye_b200", "libflye_b200.so")):
        build.build_lib()
    build.build_tools()
    if not all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", f)) for f in ("flye_restate", "trim_restate")):
        build.build_oracle()
    return True


@pytest.fixture(scope="session")
def engine(built):
    import flye_b200 as fb
    eng = fb.Engine(0)
    yield eng
    eng.close()
