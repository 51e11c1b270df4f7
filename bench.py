#!/usr/bin/env python
"""Benchmark of the reconstruction hot path on B200 (contract: see the task description / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4_heavy_hex] [--scaling weak|strong]
                    [--impl reference]

One "step" = one pass of the hot path (K1 parity tallies + K2 product/sum [+ NCCL all-reduce of K doubles])
over one batch of synthetic subexperiment results.  `value` is measured with the shot bytes already
resident in HBM; `e2e` is the same metric through ``reconstruct_expectation_values`` with host BitArrays
(one NumPy allocation per BitArray; staging + H2D + kernels + D2H inside the timed region; at N > 1 the call is
made with ``process_group=WORLD``, i.e. the sharded public API + its all-reduce).  Every measured shape carries a
same-run ``parity`` block: int64 tallies of a PUB sample bit-exact against the C oracle on the resident bytes,
and the expectation values of a bounded sub-problem against the reference's OWN source text executed on the
host (which is also what ``cpu_baseline`` times).  At N = 1 the line also carries ``per_config``: the other
BASELINE.json shapes measured the same way.  Prints exactly one JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

os.environ.setdefault("QCUT_CACHE_DIR", "/tmp/qcut_b200_cache")  # compiled-kernel cache of harness runs: not under $HOME

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "tests"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

METRIC = "reconstruct_expectation_values shots*terms/sec"
UNIT = "shot-terms/s"
DTYPE = "u8 shot rows -> int64 tallies -> f64 sums"
HEADLINE = "cfg4_heavy_hex"
PER_CONFIG = ["cfg1_tutorial", "cfg2_wire", "cfg3_wide_dense", "cfg3_wide_local", "cfg5_molecular"]
L2_BYTES = 126e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's tuples PER GPU; strong: the workload's tuples split over the GPUs")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the tuple count (debugging only)")
    ap.add_argument("--shots", type=int, default=0, help="override the shots per subexperiment (e.g. 65536 for the 115 GB strong-scaling leg)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--per-config-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per baseline measurement")
    ap.add_argument("--plan-flags", type=int, default=0, help="qcut_plan_create flags (experiments: 1 no JIT, 2 generic only, 4 staged)")
    ap.add_argument("--warm-cache", action="store_true", help="keep the on-disk kernel cache (default at N=1: a fresh, empty one)")
    ap.add_argument("--uniform-data", action="store_true", help="uniform random shot bytes instead of the biased default")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own source text (oracle/ref_loader.py) on a bounded sub-problem
# ------------------------------------------------------------------------------------------------
REF_SHOT_US, REF_TERM_US = 2.7, 0.13  # measured cost model of the reference loop: per shot, per shot and term


def reference_callable():
    """(function, kind): the reference's ``reconstruct_expectation_values`` from its own text when the sources are
    present (``oracle/_ref`` travels to the GPU box), else the oracle's literal Python port of the same loop."""
    from oracle import oracle, ref_loader

    if ref_loader.available():
        return ref_loader.load_reference()["reconstruct_expectation_values"], "reference"
    return oracle.reconstruct_literal, "port"


def size_sample(w, seconds: float) -> tuple[int, int]:
    """(tuples, shots) of a sub-problem worth ~``seconds`` of the reference's single-threaded loop."""
    per_shot_us = sum(REF_SHOT_US + REF_TERM_US * len(g.pauli_bitmasks) for c in w.collections() for g in c.groups)
    full = per_shot_us * 1e-6 * w.shots  # one tuple, all subsystems, all shots
    if 2 * full <= seconds:
        return int(min(w.n_coeffs, max(2, seconds // full))), w.shots
    c = min(2, w.n_coeffs)
    return c, int(max(64, min(w.shots, seconds / (c * per_shot_us * 1e-6))))


def sample_work(w, c: int, s: int) -> int:
    return c * s * sum(len(g.pauli_bitmasks) for col in w.collections() for g in col.groups)


def synthetic_host_results(w, c: int, s: int, seed: int) -> dict:
    """Host-only synthetic PrimitiveResults of ``c`` tuples x ``s`` shots (reference arm: no device involved)."""
    from qiskit_addon_cutting_b200.containers import BitArray, DataBin, PrimitiveResult, PubResult

    rng = np.random.default_rng(seed)
    nbq = (w.qpd_bits + 7) // 8
    out = {}
    for label, coll in zip(w.observables, w.collections()):
        pubs = []
        for _ in range(c):
            for g in coll.groups:
                nb = (g.num_measured_bits + 7) // 8
                o = rng.integers(0, 256, size=(s, nb), dtype=np.uint8) & rng.integers(0, 256, size=(s, nb), dtype=np.uint8)
                q = rng.integers(0, 256, size=(s, nbq), dtype=np.uint8)
                pubs.append(PubResult(DataBin(observable_measurements=BitArray(o, nb * 8), qpd_measurements=BitArray(q, nbq * 8))))
        out[label] = PrimitiveResult(pubs)
    return out


_REF_STATE: dict = {}  # inherited by the forked workers of the reference arm


def _reference_worker(job):
    """One process of the reference arm: the reference's function on its own sub-problem."""
    import workloads as W

    index, c, s = job
    w = _REF_STATE["workload"]
    fn = _REF_STATE["fn"]
    results = synthetic_host_results(w, c, s, w.seed + 7919 * index)
    pairs = W.coefficient_pairs(w.coefficients_global(max(c, 2))[:c])
    t0 = time.perf_counter()
    fn(results, pairs, w.observables)
    return sample_work(w, c, s), time.perf_counter() - t0


def reference_arm(args):
    """The reference's CPU implementation of the path on this box's host cores.

    ``qiskit_addon_cutting`` cannot be installed here (it needs qiskit; no wheels, no network), so the function that
    runs is the reference's own source text loaded by ``oracle/ref_loader.py`` against Qiskit stand-ins (the literal
    Python port when even the text is absent).  The reference is single-threaded; "all the host threads it can use"
    is realised the only way a user could: one process per core, each reconstructing its own tuples.
    """
    import multiprocessing as mp

    import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = W.get_workload(args.workload, args.scale)
    if args.shots:
        w.shots = args.shots
    fn, kind = reference_callable()
    for col in w.collections():  # the reference builds its own ObservableCollection per call: warm nothing, just the stand-ins
        assert col.groups
    n_procs = max(1, int(os.environ.get("QCUT_REF_PROCS", "0") or 0) or (os.cpu_count() or 1))
    # a bounded sample per step and process so that warmup + steps end within a few minutes
    # (not below 4 s: every call of the reference rebuilds its ObservableCollection, ~0.6 s on cfg4, which the full
    # problem would amortise over 97 000 tuples)
    seconds = max(4.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    _REF_STATE.update(workload=w, fn=fn)
    work = secs = 0.0
    with mp.get_context("fork").Pool(n_procs) as pool:
        # calibrate the cost model on this host with all processes busy (also the first warm-up pass)
        c0, s0 = size_sample(w, 0.5)
        t0 = time.perf_counter()
        pool.map(_reference_worker, [(i, c0, s0) for i in range(n_procs)])
        slowdown = max(0.25, (time.perf_counter() - t0) / 0.5)
        c, s = size_sample(w, seconds / slowdown)
        jobs = [(i, c, s) for i in range(n_procs)]
        for _ in range(max(0, args.warmup - 1)):
            pool.map(_reference_worker, [(i, 1, min(s, 64)) for i in range(n_procs)])
        for _ in range(args.steps):
            t0 = time.perf_counter()
            parts = pool.map(_reference_worker, jobs)
            secs += time.perf_counter() - t0  # wall clock of the slowest process
            work += sum(wk for wk, _ in parts)
    value = work / secs
    sample_desc = (f"{n_procs} processes x ({c} tuples x {s} shots of {args.workload}) per step; "
                   + ("the reference's own source text (oracle/_ref via oracle/ref_loader.py), one process per core"
                      if kind == "reference" else "python port of the reference loop, one process per core"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": dict(w.describe(), l2="n/a (cpu)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_procs, "kind": kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index: int):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(device_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [f.strip() for f in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.samples if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.samples]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


def peaks() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(name: str, scale: float):
    """dram read + write bytes of one qcut_tally_launch from the committed ``ncu --set full`` capture of this workload
    (profiles/*_k1_traffic.json); not measured in this run, hence ``traffic_source``."""
    best = None
    for path in sorted((ROOT / "profiles").glob("*_k1_traffic*.json")):
        rec = json.loads(path.read_text())
        if rec.get("workload") == name and abs(rec.get("scale", 1.0) - scale) < 1e-9:
            best = (rec["traffic_bytes_per_k1_launch"], f"profiles/{path.name} (ncu --set full capture, not this run)")
    return best or (None, None)


def host_copy_ceiling(gib: float = 1.0, array_kib: int = 64) -> dict:
    """STREAM-style copy ceiling of this host through the library's own gather (scripts/host_bw.py): ``gib`` GiB held
    in separate ``array_kib`` KiB allocations copied into one buffer with all cores.  Staging does exactly this copy
    (BitArrays -> pinned ring), so the summed e2e rate of the ranks of one host cannot exceed it."""
    from qiskit_addon_cutting_b200 import _lib

    lib = _lib.load()
    size = array_kib << 10
    n = int(gib * (1 << 30)) // size
    arrays = [np.full(size, i & 0xFF, dtype=np.uint8) for i in range(n)]
    src = np.array([a.ctypes.data for a in arrays], dtype=np.uint64)
    nbytes = np.full(n, size, dtype=np.uint64)
    offs = np.arange(n, dtype=np.uint64) * np.uint64(size)
    dst = np.zeros(n * size, dtype=np.uint8)
    threads = min(os.cpu_count() or 1, 32)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        assert lib.qcut_gather_host(src.ctypes.data, nbytes.ctypes.data, offs.ctypes.data, n, dst.ctypes.data, threads) == 0
        best = max(best, n * size / (time.perf_counter() - t0) / 1e9)
    return {"copy_gbs": best, "threads": threads, "cpus": os.cpu_count(),
            "what": f"{gib:g} GiB in {array_kib} KiB allocations copied into one buffer (bytes copied per second; read + write)"}


def tally_parity(tables, data, tallies, *, max_work: float = 3e9, max_pubs: int = 256) -> dict:
    """A spread sample of PUBs recomputed by the C oracle on the SAME resident bytes: int64 tallies must be bit-exact.
    Also reports how far the reference's sequential ``+= (1/S) * sign`` arithmetic (restated in C, full shot counts)
    is from ``tally / S`` on those PUBs -- the reference's own rounding noise at this shot count."""
    from oracle import c_oracle

    n = len(tables.pubs)
    if n == 0:
        return {"pubs_checked": 0, "tallies_bit_exact": True}
    T_of = tables.groups["n_terms"][tables.pubs["group"]].astype(np.int64)
    mean_work = float((tables.pubs["shots"].astype(np.int64) * T_of).mean())
    count = int(max(2, min(max_pubs, n, max_work // max(mean_work, 1.0))))
    picks = np.unique(np.linspace(0, n - 1, num=count, dtype=np.int64))
    sub = tables.pubs[picks].copy()
    nb_obs = tables.groups["nb_obs"][sub["group"]].astype(np.int64)
    chunks, offset, t_off = [], 0, 0
    for j in range(len(sub)):
        S, nbo, nbq = int(sub["shots"][j]), int(nb_obs[j]), int(sub["nb_qpd"][j])
        o = data[int(sub["obs_offset"][j]): int(sub["obs_offset"][j]) + S * nbo].cpu().numpy()
        q = data[int(sub["qpd_offset"][j]): int(sub["qpd_offset"][j]) + S * nbq].cpu().numpy()
        sub["obs_offset"][j] = offset
        offset += (len(o) + 15) // 16 * 16
        sub["qpd_offset"][j] = offset
        offset += (len(q) + 15) // 16 * 16
        chunks += [o, np.zeros((-len(o)) % 16, np.uint8), q, np.zeros((-len(q)) % 16, np.uint8)]
        sub["tally_offset"][j] = t_off
        t_off += int(T_of[picks[j]])
    host = np.concatenate(chunks + [np.zeros(128, np.uint8)])
    e_seq, want = c_oracle.pubs(host, sub, tables.groups, tables.masks, t_off, want_seq=True, n_threads=0)
    got = np.concatenate([tallies[int(tables.pubs["tally_offset"][p]): int(tables.pubs["tally_offset"][p]) + int(T_of[p])]
                          .cpu().numpy() for p in picks])
    shots_rep = np.repeat(sub["shots"].astype(np.float64), T_of[picks])
    exact_e = np.divide(want, shots_rep, out=np.zeros_like(shots_rep), where=shots_rep > 0)
    return {"pubs_checked": int(len(picks)), "tallies_checked": int(t_off), "tallies_bit_exact": bool(np.array_equal(got, want)),
            "checker": "oracle/oracle_seq.c on the resident bytes of a spread sample of PUBs (full shot counts)",
            "per_term_max_abs_dev_reference_arithmetic_vs_exact": float(np.max(np.abs(e_seq - exact_e))) if t_off else 0.0,
            "per_term_max_abs_expval": float(np.max(np.abs(exact_e))) if t_off else 0.0}


def measure(args, name: str, *, steps: int, warmup: int, e2e_steps: int, cpu_seconds: float, want_cpu: bool,
            want_e2e: bool, sample_clocks: bool) -> dict:
    """Measure one workload: resident throughput + K1 roofline, same-run parity, end to end, CPU baseline."""
    import torch
    import torch.distributed as dist

    import workloads as W
    from oracle import oracle
    from qiskit_addon_cutting_b200 import engine, reconstruct_expectation_values

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(warmup, 3)

    w_global = W.get_workload(name, args.scale)
    if args.shots:
        w_global.shots = args.shots
    w, coeffs_global, (c_lo, c_hi) = W.shard(w_global, rank, world, args.scaling)
    tables = W.build_tables(w, coeffs=coeffs_global[c_lo:c_hi])
    data = W.device_data(tables, w.seed * 1000 + rank, biased=not args.uniform_data)

    # ---- cold start (N = 1): the FIRST public call of this process for this shape, kernel cache empty ----------
    # (bench.py points QCUT_CACHE_DIR at a fresh directory).  The call does not wait for NVRTC: its plan compiles
    # on a background thread and the ahead-of-time kernels serve it; calls that stage < 64 MiB never compile.
    results = cold = None
    if want_e2e and world == 1:
        results = W.host_results(w, tables, W.device_fetch(data))
        pairs = W.coefficient_pairs(tables.coeffs)
        torch.cuda.synchronize()
        t = time.perf_counter()
        first = reconstruct_expectation_values(results, pairs, w.observables)
        first_ms = 1e3 * (time.perf_counter() - t)
        plans = [p for p in engine._plan_cache.values() if p.groups.tobytes() == tables.groups.tobytes()]
        state = plans[-1].jit_state() if plans else None
        for p in plans:
            p.wait()
        ready_s = time.perf_counter() - t
        cold = {"first_call_ms": first_ms, "jit_state_after_first_call": state,
                "specialised_kernels_ready_after_s": ready_s if state in (1, 2) else None,
                "kernel_cache": "empty (fresh QCUT_CACHE_DIR)" if os.environ.get("QCUT_BENCH_COLD_CACHE") else "as found",
                "note": "jit_state 1 = NVRTC still running in the background when the call returned (served by the "
                        "ahead-of-time kernels), 2 = ready, 0 = launch-bound call, never compiles"}
        cold_first = np.array(first)
    t_plan = time.perf_counter()
    batch = engine.DeviceBatch(tables, data, flags=args.plan_flags)
    plan_s = time.perf_counter() - t_plan
    out = batch.out[: tables.n_obs]
    fits_l2 = tables.data_bytes <= 2 * L2_BYTES
    flush = torch.empty(int(3 * L2_BYTES), dtype=torch.uint8, device="cuda") if fits_l2 else None

    def step():
        n = batch.launch_tally()
        n += batch.launch_reduce()
        if world > 1:
            dist.all_reduce(out)
        return n

    for _ in range(warmup):
        launches = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    sampler = ClockSampler(local) if (rank == 0 and sample_clocks) else None
    t0 = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(steps):
        if flush is not None:
            flush.fill_(k & 0xFF)  # inputs fit in L2: evict them between timed iterations
        ev[k][0].record()
        batch.launch_tally()
        ev[k][1].record()
        batch.launch_reduce()
        if world > 1:
            dist.all_reduce(out)
        ev[k][2].record()
    end.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    if flush is None:
        total_ms = start.elapsed_time(end)  # back-to-back steps, inputs far larger than L2
    else:
        total_ms = float(sum(e[0].elapsed_time(e[2]) for e in ev))  # the flushes are not part of a step
    elapsed = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    k1 = torch.tensor([np.mean([e[0].elapsed_time(e[1]) for e in ev])], dtype=torch.float64, device="cuda")
    k2_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    work = torch.tensor([float(tables.work)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_per_step = float(elapsed.item()) / steps
    k1_ms = float(k1.item())
    work_step = float(work.item())
    value = work_step / (ms_per_step * 1e-3)
    resident_out = out.cpu().numpy().copy()
    del flush

    # ---- roofline of the dominant kernel (K1): algorithmic bytes of one launch / its mean duration ----
    nbq = tables.pubs["nb_qpd"].astype(np.int64)
    nbo = tables.groups["nb_obs"][tables.pubs["group"]].astype(np.int64)
    k1_bytes = int((tables.pubs["shots"].astype(np.int64) * (nbo + nbq)).sum()) + 8 * tables.n_tallies
    peak, peak_src = peaks()
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    traffic, traffic_src = committed_traffic(name, args.scale) if (world == 1 and not args.shots) else (None, None)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "K1 parity tally (all launches of qcut_tally_launch)",
                "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": k1_ms, "peak_source": peak_src,
                "k1_share_of_step": k1_ms / ms_per_step, "k2_plus_allreduce_ms": k2_ms,
                "k1_kernels_per_launch": int(batch.tally_launches)}

    # ---- opt-in reference-rounding mode on the full resident data (shot-ordered fp64 chains + sequential K2) ----
    literal = None
    if rank == 0 and world == 1:
        lit = engine.DeviceBatch(tables, data, flags=args.plan_flags, rounding="reference")
        lit.run()
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        lit_out = lit.run()
        l1.record()
        torch.cuda.synchronize()
        lit_ms = l0.elapsed_time(l1)
        literal = {"ms_per_step": lit_ms, "value": work_step / (lit_ms * 1e-3), "unit": UNIT,
                   "max_abs_dev_vs_default_mode": float(np.max(np.abs(lit_out.cpu().numpy() - resident_out))),
                   "what": "rounding='reference' on the same resident data: every (PUB, term) chain accumulated shot by shot in "
                           "fp64 like cutting_reconstruction.py:159, tuples summed strictly in order (:168)"}
        del lit, lit_out

    # ---- same-run parity, part 1: int64 tallies of a PUB sample against the C oracle (rank 0) ----------
    parity = None
    if rank == 0:
        batch.launch_tally()
        torch.cuda.synchronize()
        parity = tally_parity(tables, data, batch.tallies)

    # ---- CPU sub-problem: reference text vs our public API vs the exact oracle (rank 0, N = 1) ----------
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        fn, kind = reference_callable()
        c, s = size_sample(w, cpu_seconds)
        sub_results = W.host_results(w, tables, W.device_fetch(data), max_coeffs=c, max_shots=s)
        sub_pairs = W.coefficient_pairs(tables.coeffs[:c])
        t = time.perf_counter()
        ref_vals = np.array(fn(sub_results, sub_pairs, w.observables))
        secs = time.perf_counter() - t
        ours_vals = np.array(reconstruct_expectation_values(sub_results, sub_pairs, w.observables))
        lit_vals = np.array(reconstruct_expectation_values(sub_results, sub_pairs, w.observables, rounding="reference"))
        exact_vals = np.array(oracle.reconstruct_exact(sub_results, sub_pairs, w.observables))
        kappa_sub = float(np.abs(tables.coeffs[:c]).sum())
        parity.update({
            "expvals_checker": f"{kind}: " + ("the reference's own source text (oracle/ref_loader.py)" if kind == "reference"
                                               else "oracle.reconstruct_literal (python port)"),
            "expvals_sample": f"first {c} tuples x first {s} shots of the resident data, all subsystems, {len(ours_vals)} observables",
            "sample_kappa": kappa_sub,
            "max_abs_dev_vs_reference": float(np.max(np.abs(ours_vals - ref_vals))),
            "max_abs_dev_vs_exact": float(np.max(np.abs(ours_vals - exact_vals))),
            "reference_vs_exact": float(np.max(np.abs(ref_vals - exact_vals))),
            "bit_identical_to_exact_oracle": bool(np.array_equal(ours_vals, exact_vals)),
            "within_1e-12_of_reference": bool(np.max(np.abs(ours_vals - ref_vals)) <= 1e-12),
            "reference_rounding_mode_bit_identical_to_reference": bool(np.array_equal(lit_vals, ref_vals)),
        })
        cpu = {"value": sample_work(w, c, s) / secs, "unit": UNIT, "cores": 1, "kind": kind, "seconds": secs,
               "sample": f"{c} tuples x {s} shots of the same resident data through "
                         + ("the reference's reconstruct_expectation_values (its own text; single-threaded, as the reference runs)"
                            if kind == "reference" else "the python port of the reference loop (single-threaded)")}
        del sub_results

    # ---- end to end through the public API with host buffers -------------------------------------------
    e2e = None
    if want_e2e:
        total_c = len(coeffs_global)
        if results is None:
            results = W.host_results(w, tables, W.device_fetch(data), total_coeffs=total_c if world > 1 else None, first_coeff=c_lo)
        if world > 1:
            pairs = W.coefficient_pairs(coeffs_global)
            group = dist.group.WORLD
        else:
            pairs = W.coefficient_pairs(tables.coeffs)
            group = None
        ceiling = host_copy_ceiling() if rank == 0 else None
        t = time.perf_counter()
        got = reconstruct_expectation_values(results, pairs, w.observables, process_group=group)  # warm-up (pinned buffer, plan cache)
        torch.cuda.synchronize()
        first = time.perf_counter() - t
        if world == 1 and first < 0.02:
            # a call of a millisecond or less: average over enough calls (~0.3 s) for a stable figure
            got = reconstruct_expectation_values(results, pairs, w.observables)
            t = time.perf_counter()
            got = reconstruct_expectation_values(results, pairs, w.observables)
            e2e_steps = int(min(300, max(e2e_steps, 0.3 / max(time.perf_counter() - t, 1e-5))))
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(e2e_steps):
            got = reconstruct_expectation_values(results, pairs, w.observables, process_group=group)
        secs_t = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(secs_t, op=dist.ReduceOp.MAX)
        e2e_secs = float(secs_t.item())
        h2d = int(tables.data_bytes + tables.pubs.nbytes)
        e2e = {"value": work_step * e2e_steps / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(8 * tables.n_obs), "ms_per_step": 1e3 * e2e_secs / e2e_steps, "steps": e2e_steps,
               "host_gbs_per_rank": h2d / (e2e_secs / e2e_steps) / 1e9,
               "host_buffers": f"{2 * len(tables.pubs)} separate NumPy allocations per rank (one per BitArray)",
               "api": "reconstruct_expectation_values(results, coefficients, observables"
                      + (", process_group=WORLD): sharded public API + NCCL all-reduce" if world > 1 else ")"),
               "host_copy_ceiling": ceiling,
               "host_copy_frac": (world * h2d / (e2e_secs / e2e_steps) / 1e9 / ceiling["copy_gbs"]) if ceiling else None,
               "staging_path": engine.last_stage_trace.get("path"),
               "dedup_ratio": engine.last_stage_trace.get("dedup", {}).get("dedup_ratio")}
        # same bytes, same path: the API result must equal the resident-data result
        if world == 1:
            assert np.array_equal(np.array(got), resident_out), "e2e result differs from resident result"
            assert np.array_equal(cold_first, resident_out), "first (cold) call differs from resident result"
        else:
            kappa = float(np.abs(coeffs_global).sum())
            assert np.max(np.abs(np.array(got) - resident_out)) <= 1e-12 * kappa, "e2e result differs from resident result"
        del results

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic" + ("" if args.uniform_data else " (per-bit Bernoulli bias 0.25/0.75, sparse qpd bits)"),
        "config": dict(w_global.describe(), l2="flushed_between_steps" if fits_l2 else "inputs_exceed_l2",
                       bytes_per_gpu=int(tables.data_bytes), work_per_gpu=int(tables.work),
                       coefficient_tuples_total=int(len(coeffs_global)), coefficient_tuples_this_rank=int(len(tables.coeffs))),
        "roofline": roofline, "parity": parity, "reference_rounding_mode": literal, "cpu_baseline": cpu, "e2e": e2e,
        "cold_start": dict(cold, synchronous_plan_s=plan_s) if cold else {"synchronous_plan_s": plan_s},
        "gpu_launches": int(launches) * steps, "gpu_launches_per_step": int(launches), "clocks": clocks,
        "checksum": float(np.sum(resident_out)),
    }
    del batch, data
    torch.cuda.empty_cache()
    return line


def measure_dedup(args) -> dict:
    """SURVEY section 8f row 4 on a wire-cutting shape: cfg2 (3 subsystems in a chain, 2 cut wires with 8 basis elements
    each).  The circuit of an end subsystem depends on ONE cut's map id, that of the middle subsystem on both, so of
    the 512 sampled tuples only 8 / 64 / 8 subexperiment sets are distinct per subsystem -- the reference still lists
    one per tuple (cutting_experiments.py:146-157).  A caller who runs each distinct circuit once passes the same
    PubResult objects many times; the engine stages and tallies each distinct PUB once."""
    import torch

    import workloads as W
    from qiskit_addon_cutting_b200 import engine, reconstruct_expectation_values
    from qiskit_addon_cutting_b200.containers import PrimitiveResult

    w = W.get_workload("cfg2_wire", 1.0)
    tables = W.build_tables(w)
    data = W.device_data(tables, w.seed * 1000, biased=not args.uniform_data)
    results = W.host_results(w, tables, W.device_fetch(data))
    C = w.n_coeffs
    shared = {}
    for (label, res), distinct in zip(results.items(), (8, 64, 8)):
        G = len(res) // C
        pubs = list(res)
        shared[label] = PrimitiveResult([pubs[((i * 8 // C) if distinct == 8 else (i * 64 // C)) % distinct * G + g]
                                         for i in range(C) for g in range(G)])
    pairs = W.coefficient_pairs(tables.coeffs)
    out = {}
    for name, flag in (("with_dedup", True), ("without_dedup", False)):
        engine._DEDUP = flag
        got = reconstruct_expectation_values(shared, pairs, w.observables)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            got = reconstruct_expectation_values(shared, pairs, w.observables)
        ms = 1e2 * (time.perf_counter() - t)
        d = engine.last_stage_trace["dedup"]
        out[name] = {"ms_per_call": ms, "h2d_shot_bytes": d["staged_shot_bytes"], "pubs_tallied": d["distinct_pubs"],
                     "result": [float(v).hex() for v in got]}
    engine._DEDUP = True
    same = out["with_dedup"].pop("result") == out["without_dedup"].pop("result")
    return {"workload": "cfg2_wire with shared result objects (8/64/8 distinct subexperiment sets for 512 tuples)",
            "pubs": int(len(tables.pubs)), "dedup_ratio": out["without_dedup"]["h2d_shot_bytes"] / out["with_dedup"]["h2d_shot_bytes"],
            "results_bit_identical": bool(same), **out}


def ours_arm(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    entry.build_library()  # serialised across ranks by a file lock; a no-op when the library is current
    if world == 1 and not args.warm_cache:
        import tempfile

        os.environ["QCUT_CACHE_DIR"] = tempfile.mkdtemp(prefix="qcut_bench_cold_")  # every plan of this run is a cold start
        os.environ["QCUT_BENCH_COLD_CACHE"] = "1"
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout at communicator creation; the contract is ONE JSON line on
        # stdout, so stdout is pointed at stderr until the first collective has run.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    want_cpu = not args.no_cpu_baseline
    with_per_config = world == 1 and not args.no_per_config and args.workload == HEADLINE and args.scale == 1.0 and not args.shots
    per_by_name = {}

    def per_config(names):
        for name in names:
            t0 = time.perf_counter()
            try:
                rec = measure(args, name, steps=args.per_config_steps, warmup=3, e2e_steps=max(3, args.e2e_steps),
                              cpu_seconds=min(args.cpu_seconds, 6.0), want_cpu=want_cpu, want_e2e=not args.no_e2e,
                              sample_clocks=False)
                keep = ("value", "unit", "ms_per_step", "steps", "config", "roofline", "parity", "reference_rounding_mode", "cpu_baseline", "e2e", "cold_start",
                        "gpu_launches_per_step", "checksum")
                rec = {k: rec[k] for k in keep}
            except Exception as exc:  # a failing shape must not take the headline line down with it
                rec = {"config": {"workload": name}, "error": repr(exc)}
            rec["wall_s"] = time.perf_counter() - t0
            per_by_name[name] = rec

    # The sub-millisecond shapes are measured first, in a process that has not yet allocated and freed the 388 000
    # BitArrays of the headline shape (their per-call host time was 2x higher when measured after it).
    small = [n for n in PER_CONFIG if n in ("cfg1_tutorial", "cfg2_wire")]
    if with_per_config:
        per_config(small)
    line = measure(args, args.workload, steps=args.steps, warmup=args.warmup, e2e_steps=args.e2e_steps,
                   cpu_seconds=args.cpu_seconds, want_cpu=want_cpu, want_e2e=not args.no_e2e, sample_clocks=True)
    if with_per_config:
        per_config([n for n in PER_CONFIG if n not in small])
        line["per_config"] = [per_by_name[n] for n in PER_CONFIG]
        try:
            line["dedup"] = measure_dedup(args)
        except Exception as exc:
            line["dedup"] = {"error": repr(exc)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
