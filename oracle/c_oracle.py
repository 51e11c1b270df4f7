"""ctypes wrapper of oracle_seq.c (TEST INFRASTRUCTURE ONLY, see oracle/oracle.py)."""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle_seq.so"
_lib = None


def build() -> Path:
    src = HERE / "oracle_seq.c"
    if not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        lib.oracle_pub.restype = None
        lib.oracle_pub.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.oracle_pubs.restype = None
        lib.oracle_pubs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def pub(obs: np.ndarray, qpd: np.ndarray, mask_bytes: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(e_seq, tally) of one PUB; ``mask_bytes`` is (T, nb_obs) uint8 big-endian rows."""
    shots = qpd.shape[0]
    obs = np.ascontiguousarray(obs[:shots], dtype=np.uint8)
    qpd = np.ascontiguousarray(qpd, dtype=np.uint8)
    mask_bytes = np.ascontiguousarray(mask_bytes, dtype=np.uint8)
    T = mask_bytes.shape[0]
    e = np.zeros(T, dtype=np.float64)
    tally = np.zeros(T, dtype=np.int64)
    load().oracle_pub(obs.ctypes.data, qpd.ctypes.data, shots, obs.shape[1], qpd.shape[1], mask_bytes.ctypes.data, T,
                      e.ctypes.data, tally.ctypes.data)
    return e, tally


def pubs(data: np.ndarray, pub_table: np.ndarray, group_table: np.ndarray, masks: np.ndarray, n_tallies: int,
         *, want_seq: bool = True, n_threads: int = 0) -> tuple[np.ndarray | None, np.ndarray]:
    """All PUBs of a staged buffer (same tables as the product's C ABI)."""
    e = np.zeros(max(n_tallies, 1), dtype=np.float64) if want_seq else None
    tally = np.zeros(max(n_tallies, 1), dtype=np.int64)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    load().oracle_pubs(data.ctypes.data, pub_table.ctypes.data, len(pub_table), group_table.ctypes.data,
                       masks.ctypes.data if masks.size else None, e.ctypes.data if want_seq else None,
                       tally.ctypes.data, n_threads)
    return (e[:n_tallies] if want_seq else None), tally[:n_tallies]


def max_threads() -> int:
    return load().oracle_max_threads()
