"""Host memory-copy ceiling of a box (STREAM-style copy), through the library's own gather (qcut_gather_host).

    python scripts/host_bw.py [--gib 4] [--array-kib 64]

Copies `gib` GiB held in separate `array-kib` KiB NumPy allocations (the shape of a PrimitiveResult's BitArrays)
into one destination buffer with 1, 2, 4, ... threads and prints GB/s of bytes COPIED (each byte is one DRAM read
and one DRAM write).  The end-to-end staging path does this copy into a pinned ring and the DMA engine reads the
ring once more, so a host whose copy ceiling is X GB/s cannot stage faster than ~X GB/s summed over its ranks
unless the ring stays in the caches."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from qiskit_addon_cutting_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gib", type=float, default=4.0)
ap.add_argument("--array-kib", type=int, default=64)
args = ap.parse_args()
lib = _lib.load()
n = int(args.gib * (1 << 30)) // (args.array_kib << 10)
size = args.array_kib << 10
arrays = [np.full(size, i & 0xFF, dtype=np.uint8) for i in range(n)]
src = np.array([a.ctypes.data for a in arrays], dtype=np.uint64)
nbytes = np.full(n, size, dtype=np.uint64)
dst_off = (np.arange(n, dtype=np.uint64) * np.uint64(size))
dst = np.empty(n * size, dtype=np.uint8)
dst[:] = 0
out = {"bytes": int(n * size), "array_kib": args.array_kib, "cpus": os.cpu_count(), "copy_gbs": {}}
t = 1
while t <= (os.cpu_count() or 1):
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        rc = lib.qcut_gather_host(src.ctypes.data, nbytes.ctypes.data, dst_off.ctypes.data, n, dst.ctypes.data, t)
        assert rc == 0
        best = max(best, n * size / (time.perf_counter() - t0) / 1e9)
    out["copy_gbs"][str(t)] = round(best, 1)
    t *= 2
print(json.dumps(out))
