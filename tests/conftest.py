import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib_built():
    """Build libagt.so in-tree if needed (nvcc cross-compiles without a GPU)."""
    from accurate_aprilgroup_tracking_b200 import _build
    return _build.build()


@pytest.fixture(scope="session")
def ctx1080(lib_built):
    from accurate_aprilgroup_tracking_b200 import synth
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    c = AgtContext(0, synth.CAMERA_1080P.mtx, None)
    c.set_synthetic_model()
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctxvga(lib_built):
    from accurate_aprilgroup_tracking_b200 import synth
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    c = AgtContext(0, synth.CAMERA_VGA.mtx, None)
    c.set_synthetic_model()
    yield c
    c.close()
