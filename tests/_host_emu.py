"""Compile the NVRTC translation unit of a plan with g++ (QCUT_HOST_EMU) and run it lane by lane.

The generated source is the same text the GPU gets; only the loads and the two CUDA intrinsics it
uses are shimmed.  A warp is emulated as 32 sequential lanes whose packed counters are summed --
which is exactly what the device-side flush (REDUX + integer atomics) does.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from qiskit_addon_cutting_b200 import _lib

BUILD = Path(tempfile.gettempdir()) / "qcut_b200_emu_build"  # outside the repo: built artefacts do not travel

SHIM = r"""
#include <cstdint>
#include <cstring>
#define QCUT_HOST_EMU 1
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t both = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((both >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
"""


def compile_emulator(source: str) -> C.CDLL:
    BUILD.mkdir(exist_ok=True)
    text = SHIM + source
    tag = hashlib.sha1(text.encode()).hexdigest()[:16]
    so = BUILD / f"emu_{tag}.so"
    if not so.exists():
        src = BUILD / f"emu_{tag}.cpp"
        src.write_text(text)
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-o", str(so), str(src)], check=True)
    lib = C.CDLL(str(so))
    entries = {  # which of them a source exports depends on the kind of kernel
        "qcut_emu_lane_tile": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p],
        "qcut_emu_lane_stream": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p],
        "qcut_emu_rt_lane_tile": [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
        "qcut_emu_lane_tile_variant": [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p],  # switch kernels
    }
    for name, argtypes in entries.items():
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = argtypes
    return lib


def mask_words(tables, g: int) -> np.ndarray:
    """Little-endian 32-bit words of group g's mask rows in memory order (what qcut_plan_create uploads)."""
    grp = tables.groups[g]
    T, nb = int(grp["n_terms"]), int(grp["nb_obs"])
    rows = tables.masks[int(grp["mask_offset"]): int(grp["mask_offset"]) + T * nb].reshape(T, nb)
    padded = np.zeros((T, (nb + 3) // 4 * 4), dtype=np.uint8)
    padded[:, :nb] = rows
    return np.ascontiguousarray(padded).view("<u4").reshape(T, -1).copy()


def emulate_tallies(tables, data: np.ndarray, engine: str = "jit") -> tuple[np.ndarray, np.ndarray]:
    """Tallies of every PUB whose group the bit-sliced kernels cover, emulated lane by lane.

    engine = "jit": the group's NVRTC translation unit (masks baked in); "rt": run-time-mask engine.
    Returns (tallies, covered mask per PUB).
    """
    tallies = np.zeros(max(tables.n_tallies, 1), dtype=np.int64)
    sources = {g: _lib.generate_source(tables.groups, tables.masks, g) for g in range(len(tables.groups))}
    if engine == "rt":  # the run-time-mask kernel covers rows of up to 8 bytes
        sources = {g: (src if int(tables.groups[g]["nb_obs"]) <= 8 else "") for g, src in sources.items()}
    covered = np.array([bool(sources[int(g)]) for g in tables.pubs["group"]], dtype=bool)
    emus = {g: compile_emulator(src) for g, src in sources.items() if src}
    odd = np.zeros(4096, dtype=np.uint32)
    for p, pub in enumerate(tables.pubs):
        g = int(pub["group"])
        if g not in emus:
            continue
        emu = emus[g]
        grp = tables.groups[g]
        nb, nbq, shots, T = int(grp["nb_obs"]), int(pub["nb_qpd"]), int(pub["shots"]), int(grp["n_terms"])
        mw = mask_words(tables, g)
        for shot0 in range(0, shots, _lib.TILE_SHOTS):
            ns = min(_lib.TILE_SHOTS, shots - shot0)
            obs_ptr = data.ctypes.data + int(pub["obs_offset"]) + shot0 * nb
            qpd_ptr = data.ctypes.data + int(pub["qpd_offset"]) + shot0 * nbq
            total = np.zeros(T, dtype=np.int64)
            for lane in range(32):
                if engine == "jit":
                    n = emu.qcut_emu_lane_tile(obs_ptr, qpd_ptr, nbq, ns, lane, odd.ctypes.data)
                else:
                    n = emu.qcut_emu_rt_lane_tile(nb, obs_ptr, qpd_ptr, nbq, ns, lane, mw.ctypes.data, mw.shape[1], T, odd.ctypes.data)
                assert n == T
                total += odd[:T]
            tallies[int(pub["tally_offset"]): int(pub["tally_offset"]) + T] += ns - 2 * total
    return tallies[: tables.n_tallies], covered
