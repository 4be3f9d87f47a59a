import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    # the oracle is test infrastructure: build it on demand (seconds, g++ only)
    from oracle import bridge

    bridge.build()
    yield


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a GPU box.  No skip: a gpu-marked test without a device fails."""
    from plainmp_b200 import _capi

    lib = _capi.load()
    assert lib.pmp_device_count() >= 1, "gpu-marked test needs a CUDA device"
    return lib
