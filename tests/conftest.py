import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
# compiled-kernel cache of the test runs: not under $HOME
os.environ.setdefault("QCUT_CACHE_DIR", "/tmp/qcut_b200_cache")
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry

    entry.build_library()
    from qiskit_addon_cutting_b200 import _lib

    return _lib.load()
