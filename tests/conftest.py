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


def pytest_collection_modifyitems(config, items):
    # a kernel that never returns would otherwise hold the GPU box until the caller's own limit:
    # every gpu-marked test gets a hard limit enforced from a watchdog thread (a blocked CUDA
    # synchronise inside a ctypes call never sees a signal), which ends the process with a stack dump
    if not config.pluginmanager.hasplugin("timeout"):
        # pytest-timeout is a test dependency of the GPU tier (tests/requirements.txt); without it the tests still run,
        # but nothing bounds a kernel that never returns: say so instead of skipping the watchdog silently
        import warnings

        if any(item.get_closest_marker("gpu") for item in items):
            warnings.warn("pytest-timeout is not installed: gpu-marked tests run without the 900 s watchdog", RuntimeWarning)
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))
