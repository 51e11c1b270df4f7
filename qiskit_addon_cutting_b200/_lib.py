"""ctypes binding of ``libqcut_b200.so`` (the C ABI declared in ``include/qcut_b200.h``).

There is deliberately no fallback: if the shared library is missing the import of the
compute path fails with instructions to build it, and if it is present but no CUDA device
is visible every compute entry point raises.  The CPU oracle under ``oracle/`` is test
infrastructure and is never consulted from here.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PACKAGE_DIR = Path(__file__).resolve().parent
LIB_PATH = PACKAGE_DIR / "libqcut_b200.so"

TILE_SHOTS = 4096
MAX_OBS_BYTES = 64
DATA_SLACK = 128  # readable bytes required after the last matrix (QCUT_DATA_SLACK)
ABI_VERSION = 4
REDUCE_SINGLE_SLOT, REDUCE_SEQUENTIAL = 1, 2
PLAN_NO_JIT = 1
PLAN_GENERIC_ONLY = 2
PLAN_STAGED = 4
PLAN_ASYNC_JIT = 8
KERNEL_GENERIC, KERNEL_SLICED_RT, KERNEL_JIT0 = 0, 1, 2


class QcutError(RuntimeError):
    """A call into libqcut_b200 failed."""


class GroupDesc(C.Structure):
    _fields_ = [("n_terms", C.c_uint32), ("nb_obs", C.c_uint32), ("mask_offset", C.c_uint32), ("reserved", C.c_uint32)]


class PubDesc(C.Structure):
    _fields_ = [
        ("obs_offset", C.c_uint64),
        ("qpd_offset", C.c_uint64),
        ("tally_offset", C.c_uint64),
        ("shots", C.c_uint32),
        ("group", C.c_uint32),
        ("nb_qpd", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class Slot(C.Structure):
    _fields_ = [("group", C.c_uint32), ("term", C.c_uint32)]


class LabelDesc(C.Structure):
    _fields_ = [("pub_base", C.c_uint64), ("num_groups", C.c_uint32), ("lookup_base", C.c_uint32)]


# numpy views of the same structs (used to build tables without Python-level loops)
GROUP_DTYPE = np.dtype([("n_terms", "<u4"), ("nb_obs", "<u4"), ("mask_offset", "<u4"), ("reserved", "<u4")])
PUB_DTYPE = np.dtype(
    [
        ("obs_offset", "<u8"),
        ("qpd_offset", "<u8"),
        ("tally_offset", "<u8"),
        ("shots", "<u4"),
        ("group", "<u4"),
        ("nb_qpd", "<u4"),
        ("reserved", "<u4"),
    ]
)
SLOT_DTYPE = np.dtype([("group", "<u4"), ("term", "<u4")])
LABEL_DTYPE = np.dtype([("pub_base", "<u8"), ("num_groups", "<u4"), ("lookup_base", "<u4")])
assert GROUP_DTYPE.itemsize == C.sizeof(GroupDesc) == 16
assert PUB_DTYPE.itemsize == C.sizeof(PubDesc) == 40
assert SLOT_DTYPE.itemsize == C.sizeof(Slot) == 8
assert LABEL_DTYPE.itemsize == C.sizeof(LabelDesc) == 16

_VP, _I64, _INT, _SZ = C.c_void_p, C.c_int64, C.c_int, C.c_size_t

# name -> (restype, argtypes); this table is also what the CPU tests check the .so against.
SIGNATURES = {
    "qcut_abi_version": (_INT, []),
    "qcut_last_error": (C.c_char_p, []),
    "qcut_device_count": (_INT, []),
    "qcut_plan_create": (_INT, [_VP, _INT, _VP, _SZ, C.c_uint, C.POINTER(_VP)]),
    "qcut_plan_destroy": (None, [_VP]),
    "qcut_plan_group_kernel": (_INT, [_VP, _INT]),
    "qcut_plan_source": (_SZ, [_VP, _INT, C.c_char_p, _SZ]),
    "qcut_plan_jit_state": (_INT, [_VP]),
    "qcut_plan_wait": (_INT, [_VP]),
    "qcut_schedule_create": (_INT, [_VP, _VP, _I64, C.POINTER(_VP)]),
    "qcut_schedule_destroy": (None, [_VP]),
    "qcut_schedule_num_tiles": (_I64, [_VP]),
    "qcut_schedule_device_pubs": (_VP, [_VP]),
    "qcut_tally_launch": (_INT, [_VP, _VP, _VP, _VP, _I64, _VP, C.POINTER(_INT)]),
    "qcut_reduce_workspace_bytes": (_SZ, [_I64, _INT]),
    "qcut_reduce_launch": (_INT, [_VP, _VP, _VP, _INT, _VP, _VP, _VP, _I64, _INT, _VP, _VP, _INT, _VP]),
    "qcut_expvals_literal_launch": (_INT, [_VP, _VP, _VP, _VP, _I64, _VP]),
    "qcut_tallies_to_expvals_launch": (_INT, [_VP, _VP, _I64, _VP, _VP, _VP]),
    "qcut_quasi_launch": (_INT, [_VP, _I64, _VP, _VP, _VP, _VP, _INT, _VP, _INT, _VP, _VP, _VP]),
    "qcut_reduce_expvals_launch": (_INT, [_VP, _VP, _VP, _INT, _VP, _VP, _VP, _I64, _INT, _VP, _VP, _INT, _VP]),
    "qcut_recombine_launch": (_INT, [_VP, _INT, _VP, _VP, _VP, _INT, _VP, _VP]),
    "qcut_dedup_create": (_INT, [C.POINTER(_VP)]),
    "qcut_dedup_destroy": (None, [_VP]),
    "qcut_dedup_clear": (None, [_VP]),
    "qcut_dedup_block": (_I64, [_VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP]),
    "qcut_gather_host": (_INT, [_VP, _VP, _VP, _I64, _VP, _INT]),
    "qcut_stage_h2d": (_INT, [_VP, _VP, _VP, _I64, _VP, _SZ, _VP, _INT, _VP]),
    "qcut_generate_source": (_SZ, [_VP, _INT, _VP, _SZ, _INT, C.c_char_p, _SZ]),
    "qcut_compile_source": (_INT, [C.c_char_p, C.POINTER(_SZ), _VP, _SZ]),
    "qcut_plan_compile_check": (_INT, [_VP, _INT, _VP, _SZ, C.POINTER(_INT), C.POINTER(_SZ)]),
    "qcut_plan_kernel_source": (_SZ, [_VP, _INT, _VP, _SZ, _INT, C.c_char_p, _SZ, _VP, _VP]),
}

_lib = None
_pyhelper = None


def load_pyhelper():
    """The CPython-side result walker (``csrc/pyhelper.c``), loaded with the GIL held."""
    global _pyhelper
    if _pyhelper is None:
        path = PACKAGE_DIR / "_qcut_pyhelper.so"
        if not path.exists():
            raise ImportError(f"{path} not found: run `make -C qiskit_addon_cutting_b200/csrc`.")
        lib = C.PyDLL(str(path))
        lib.qcut_py_collect.restype = C.c_int64
        lib.qcut_py_collect.argtypes = [C.py_object, C.c_int64, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, C.py_object]
        _pyhelper = lib
    return _pyhelper


def load() -> C.CDLL:
    """Load the shared library (once).  Raises with build instructions if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("QCUT_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: the CUDA library of qiskit_addon_cutting_b200 has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C qiskit_addon_cutting_b200/csrc`. There is no CPU fallback."
        )
    lib = C.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.qcut_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.qcut_abi_version()} != expected {ABI_VERSION}; rebuild.")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().qcut_last_error().decode(errors="replace")
        raise QcutError(f"{what} failed (status {rc}): {msg}")


def require_device() -> None:
    if load().qcut_device_count() < 1:
        raise QcutError(
            "qiskit_addon_cutting_b200 needs a CUDA device (built for NVIDIA B200, sm_100a); "
            "none is visible and there is no CPU fallback."
        )


def np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def generate_source(groups: np.ndarray, masks: np.ndarray, group: int) -> str:
    """CUDA source NVRTC receives when ``group`` is specialised ('' if it cannot be)."""
    lib = load()
    groups = np.ascontiguousarray(groups, dtype=GROUP_DTYPE)
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    mptr = np_ptr(masks) if masks.size else None
    need = lib.qcut_generate_source(np_ptr(groups), len(groups), mptr, masks.size, group, None, 0)
    if need <= 1:
        return ""
    buf = C.create_string_buffer(need)
    lib.qcut_generate_source(np_ptr(groups), len(groups), mptr, masks.size, group, buf, need)
    return buf.value.decode()


def compile_source(source: str) -> bytes:
    """NVRTC-compile a generated translation unit to an sm_100a cubin (no GPU needed)."""
    lib = load()
    size = _SZ(0)
    src = source.encode()
    check(lib.qcut_compile_source(src, C.byref(size), None, 0), "qcut_compile_source")
    buf = C.create_string_buffer(size.value)
    check(lib.qcut_compile_source(src, C.byref(size), buf, size.value), "qcut_compile_source")
    return buf.raw


def plan_kernel_source(groups: np.ndarray, masks: np.ndarray, kernel: int) -> tuple[str, np.ndarray, np.ndarray]:
    """(source of specialised kernel ``kernel`` of the plan these groups would get, kernel index per group, network
    index per group); the source is empty when there is no such kernel.  Host only."""
    lib = load()
    groups = np.ascontiguousarray(groups, dtype=GROUP_DTYPE)
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    mptr = np_ptr(masks) if masks.size else None
    kernel_of = np.zeros(len(groups), dtype=np.int32)
    variant_of = np.zeros(len(groups), dtype=np.int32)
    need = lib.qcut_plan_kernel_source(np_ptr(groups), len(groups), mptr, masks.size, kernel, None, 0, np_ptr(kernel_of), np_ptr(variant_of))
    if need == 0:
        return "", kernel_of, variant_of
    buf = C.create_string_buffer(need)
    lib.qcut_plan_kernel_source(np_ptr(groups), len(groups), mptr, masks.size, kernel, buf, need, None, None)
    return buf.value.decode(), kernel_of, variant_of


def plan_compile_check(groups: np.ndarray, masks: np.ndarray) -> tuple[int, int]:
    """Run every NVRTC compilation a plan for these groups needs (no GPU); returns (#kernels, cubin bytes)."""
    lib = load()
    groups = np.ascontiguousarray(groups, dtype=GROUP_DTYPE)
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    n, size = _INT(0), _SZ(0)
    check(
        lib.qcut_plan_compile_check(np_ptr(groups), len(groups), np_ptr(masks) if masks.size else None, masks.size,
                                    C.byref(n), C.byref(size)),
        "qcut_plan_compile_check",
    )
    return n.value, size.value
