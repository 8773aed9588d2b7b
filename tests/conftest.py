import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("LOGURU_LEVEL", "WARNING")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    try:
        from loguru import logger

        logger.remove()
        logger.add(sys.stderr, level="WARNING")
    except ImportError:
        pass


def _has_cuda() -> bool:
    try:
        from dynamicprogramming_b200 import _ffi

        return _ffi.lib().pi_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def ref_runner():
    """The reference's own kernels (oracle/_ref cubins).  Parity tests need them:
    a missing oracle is a failure, not a skip."""
    from oracle import ref_runner as rr

    assert rr.available(), "oracle/_ref/ is missing: run `python oracle/build_ref.py` where /root/reference exists"
    return rr
