"""Where does K1's time go?  The cfg4 kernel (NVRTC-specialised, all-at-once engine) on L2-resident rows, with pieces removed.

The tile list cycles over a few hundred PUBs (29 MB: served by L2), so HBM is out of the picture; variants are made by
editing the generated source text: no transposition, no term network, no POPC, no qpd parity plane.  Prints ms and the
HBM-equivalent rate (algorithmic bytes / time) of each.  Usage: python scripts/micro/k1_compute_only.py [out.txt]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
os.environ.setdefault("QCUT_JIT_PIPE", "0")
import workloads as W  # noqa: E402
from cuda.bindings import driver as cu  # noqa: E402
from qiskit_addon_cutting_b200 import _lib  # noqa: E402


def ck(res):
    err, *rest = res
    assert err == cu.CUresult.CUDA_SUCCESS, err
    return rest[0] if len(rest) == 1 else rest


def variant(src: str, kind: str) -> str:
    if kind == "base":
        return src
    if kind == "notranspose":
        return src.replace("QCUT_DEV void qcut_transpose32(uint32_t* a) {", "QCUT_DEV void qcut_transpose32(uint32_t* a) {\n  return;", 1)
    if kind == "nopopc":
        return src.replace("(uint32_t)__popc(x)", "(x & 0xffu)")
    if kind == "noqpd":
        old = "    qcut_qpd_loads<NB, FULL, Src>(src, n, lane, qraw);\n    qcut_qpd_planes<NB>(qraw, qp);"
        assert old in src
        return src.replace(old, "    qp[0] = (uint32_t)n;", 1)
    if kind in ("noterms", "noterms_notranspose"):
        head = src.index("  static QCUT_MDEV void run(")
        body = src.index("{\n", head) + 2
        end = src.index("\n  }\n", body)
        new = "    uint32_t x = qp;\n    for (int i = 0; i < 32; ++i) x ^= P[i] ^ Q[i];\n    acc[0] += (uint32_t)__popc(x);\n"
        out = src[:body] + new + src[end + 1:]
        return variant(out, "notranspose") if kind.endswith("notranspose") else out
    if kind == "noflush":
        old = "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);"
        assert old in src
        return src.replace(old, "    if (acc[0] == 0x12345u) tallies[0] = acc[1];")
    if kind == "noredux":
        assert "__reduce_add_sync(0xffffffffu, " in src
        return src.replace("__reduce_add_sync(0xffffffffu, ", "(")
    if kind == "nored":
        old = 'asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");'
        assert old in src
        return src.replace(old, 'if (v == 0x123456789ull) *p = v;')
    if kind == "noredux_nored":
        return variant(variant(src, "noredux"), "nored")
    raise ValueError(kind)


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    tab = W.build_tables(W.get_workload("cfg4_heavy_hex", 0.001))
    src = _lib.generate_source(tab.groups, tab.masks, 0)
    T = int(tab.groups[0]["n_terms"])
    shots, pub_bytes = 8192, 8192 * 9
    n_tiles = 2 * 97000

    def setup(n_unique):
        data = torch.randint(0, 256, (n_unique * pub_bytes + 4096,), dtype=torch.uint8, device="cuda")
        pubs = np.zeros(n_unique, dtype=_lib.PUB_DTYPE)
        pubs["obs_offset"] = np.arange(n_unique, dtype=np.uint64) * pub_bytes
        pubs["qpd_offset"] = pubs["obs_offset"] + shots * 8
        pubs["tally_offset"] = np.arange(n_unique, dtype=np.uint64) * T
        pubs["shots"], pubs["group"], pubs["nb_qpd"] = shots, 0, 1
        tiles = np.zeros((n_tiles, 2), dtype=np.uint32)
        tiles[:, 0] = (np.arange(n_tiles) // 2) % n_unique
        tiles[:, 1] = np.arange(n_tiles) % 2
        d_pubs = torch.from_numpy(pubs.view(np.uint8).copy()).cuda()
        d_tiles = torch.from_numpy(tiles.view(np.uint8).reshape(-1).copy()).cuda()
        tallies = torch.zeros(n_unique * T, dtype=torch.int64, device="cuda")
        return data, d_pubs, d_tiles, tallies

    def run(kind, bufs, label):
        cubin = _lib.compile_source(variant(src, kind))
        mod = ck(cu.cuModuleLoadData(cubin))
        fn = ck(cu.cuModuleGetFunction(mod, b"qcut_tally_sliced"))
        data, d_pubs, d_tiles, tallies = bufs
        args = (np.array([data.data_ptr()], dtype=np.uint64), np.array([d_pubs.data_ptr()], dtype=np.uint64),
                np.array([d_tiles.data_ptr()], dtype=np.uint64), np.array([n_tiles], dtype=np.int64),
                np.array([tallies.data_ptr()], dtype=np.uint64))
        argp = np.array([a.ctypes.data for a in args], dtype=np.uint64)
        stream = torch.cuda.current_stream().cuda_stream

        def launch():
            ck(cu.cuLaunchKernel(fn, 296, 1, 1, 256, 1, 1, 0, stream, argp.ctypes.data, 0))

        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = n_tiles * 4096 * 9 / ms * 1e-6
        line = f"{label:<58s} {ms:7.3f} ms  {gbs:7.0f} GB/s-equivalent  ({gbs / 6551:.3f} of 6551)"
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()
        ck(cu.cuModuleUnload(mod))

    hbm = setup(20000)
    run("base", hbm, "base, rows from HBM (1.5 GB of PUBs)")
    del hbm
    l2 = setup(400)
    for kind, label in [("base", "base, rows from L2 (29 MB of PUBs)"), ("notranspose", "no transposition"),
                        ("noterms", "no term network (transposition + XOR of all planes)"),
                        ("noterms_notranspose", "neither (loads, qpd plane, flush)"), ("nopopc", "terms without POPC (x & 0xff)"),
                        ("noqpd", "no qpd parity plane"), ("noflush", "no flush"), ("noredux", "flush without REDUX"),
                        ("nored", "flush without RED.64"), ("noredux_nored", "flush without REDUX and RED.64")]:
        run(kind, l2, label)


main()
