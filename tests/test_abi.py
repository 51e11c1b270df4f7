"""The C-ABI library loads and exports exactly what include/qcut_b200.h declares (no compute calls here)."""

import ctypes
import re
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "qcut_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcut_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from qiskit_addon_cutting_b200 import _lib

    declared = _declared()
    assert len(declared) >= 14
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in qcut_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"


def test_abi_version_and_structs(lib):
    from qiskit_addon_cutting_b200 import _lib

    assert lib.qcut_abi_version() == _lib.ABI_VERSION == 4
    header = (ROOT / "include" / "qcut_b200.h").read_text()
    assert f"#define QCUT_TILE_SHOTS {_lib.TILE_SHOTS}" in header
    assert f"#define QCUT_MAX_OBS_BYTES {_lib.MAX_OBS_BYTES}" in header
    assert f"#define QCUT_DATA_SLACK {_lib.DATA_SLACK}" in header
    assert _lib.PUB_DTYPE.itemsize == 40 and _lib.GROUP_DTYPE.itemsize == 16


def test_errors_are_reported_not_thrown(lib):
    from qiskit_addon_cutting_b200 import _lib

    plan = ctypes.c_void_p()
    assert lib.qcut_plan_create(None, 0, None, 0, 0, ctypes.byref(plan)) == -1
    assert b"no groups" in lib.qcut_last_error()
    groups = np.zeros(1, dtype=_lib.GROUP_DTYPE)
    groups[0] = (1, 65, 0, 0)  # 65-byte rows: wider than QCUT_MAX_OBS_BYTES
    masks = np.zeros(65, dtype=np.uint8)
    rc = lib.qcut_plan_create(groups.ctypes.data, 1, masks.ctypes.data, masks.size, 0, ctypes.byref(plan))
    assert rc == -1 and b"unsupported observable row width" in lib.qcut_last_error()


def test_gather_host(lib):
    rng = np.random.default_rng(0)
    arrays = [rng.integers(0, 256, size=int(n), dtype=np.uint8) for n in rng.integers(0, 5000, size=200)]
    sizes = np.array([a.nbytes for a in arrays], dtype=np.uint64)
    offsets = np.zeros(len(arrays), dtype=np.uint64)
    offsets[1:] = np.cumsum((sizes[:-1] + 15) // 16 * 16)
    src = np.array([a.ctypes.data for a in arrays], dtype=np.uint64)
    dst = np.zeros(int(offsets[-1] + sizes[-1]) + 16, dtype=np.uint8)
    for threads in (1, 4, 0):
        dst[:] = 0
        assert lib.qcut_gather_host(src.ctypes.data, sizes.ctypes.data, offsets.ctypes.data, len(arrays), dst.ctypes.data, threads) == 0
        for a, off in zip(arrays, offsets):
            assert np.array_equal(dst[int(off): int(off) + a.nbytes], a)
