import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices() -> int:
    """Devices the CUDA runtime sees, asked through the driver library (no torch import, no product code)."""
    import ctypes

    try:
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0 or cuda.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return int(n.value)
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped, not failed, on a box without a CUDA device (plain `pytest tests` stays green).
    A missing libgbp_b200.so on a GPU box is NOT a reason to skip: those tests must fail loudly there."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is test infrastructure; build it once from its committed Makefile."""
    from oracle import oracle

    oracle.build()
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
