#!/usr/bin/env python
"""Benchmark of the reconstruction hot path on B200 (contract: see the task description / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4_heavy_hex] [--impl reference]

One "step" = one pass of the hot path (K1 parity tallies + K2 product/sum [+ NCCL all-reduce of K doubles])
over one batch of synthetic subexperiment results.  `value` is measured with the shot bytes already
resident in HBM; `e2e` is the same metric through ``reconstruct_expectation_values`` with host BitArrays
(staging + H2D + kernels + D2H inside the timed region).  Prints exactly one JSON line on rank 0.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4_heavy_hex")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the tuple count (debugging only)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per baseline measurement")
    ap.add_argument("--plan-flags", type=int, default=0, help="qcut_plan_create flags (experiments: 1 no JIT, 2 generic only, 4 staged)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's loop (the reference is single-threaded pure Python)
# ------------------------------------------------------------------------------------------------
def cpu_sample(workload, tables, host_bytes_of_pub, seconds: float, count: int = 64):
    """Pick `count` (pub index, shots) pairs worth ~`seconds` of the reference's Python loop (2.7 us/shot + 0.13 us/term)."""
    sample, budget = [], seconds
    order = np.linspace(0, len(tables.pubs) - 1, num=min(len(tables.pubs), count), dtype=np.int64)
    per_pub = seconds / len(order)
    for p in order:
        pub = tables.pubs[p]
        T = int(tables.groups[pub["group"]]["n_terms"])
        shots = int(min(pub["shots"], max(64, per_pub / (2.7e-6 + 0.13e-6 * T))))
        sample.append((int(p), shots))
        budget -= shots * (2.7e-6 + 0.13e-6 * T)
    return sample


def run_reference_port(tables, sample, fetch):
    """Literal restatement of cutting_reconstruction.py:152-161,196-222 on the sample; returns (work, seconds)."""
    from oracle import oracle

    work, t0 = 0, time.perf_counter()
    for p, shots in sample:
        pub = tables.pubs[p]
        grp = tables.groups[pub["group"]]
        nb, T = int(grp["nb_obs"]), int(grp["n_terms"])
        obs, qpd = fetch(p, shots)
        rows = tables.masks[int(grp["mask_offset"]): int(grp["mask_offset"]) + T * nb].reshape(T, nb)
        masks = [int.from_bytes(r.tobytes(), "big") for r in rows]
        oracle.pub_expvals_literal(obs, qpd, masks)
        work += shots * T
    return work, time.perf_counter() - t0


def run_compiled_port(tables, host_bytes, n_threads: int):
    from oracle import c_oracle

    t0 = time.perf_counter()
    c_oracle.pubs(host_bytes, tables.pubs, tables.groups, tables.masks, tables.n_tallies, want_seq=True, n_threads=n_threads)
    return time.perf_counter() - t0


def synthetic_fetch(workload, tables, seed):
    """Host-side synthetic shot rows for the CPU arm (same distribution as the device data: uniform bytes)."""
    rng = np.random.default_rng(seed)

    def fetch(p, shots):
        pub = tables.pubs[p]
        nb = int(tables.groups[pub["group"]]["nb_obs"])
        return (rng.integers(0, 256, size=(shots, nb), dtype=np.uint8),
                rng.integers(0, 256, size=(shots, int(pub["nb_qpd"])), dtype=np.uint8))

    return fetch


_REF_STATE: dict = {}  # inherited by the forked workers of the reference arm


def _reference_worker(job):
    """One process of the reference arm: the python port of the reference loop on its share of the sample."""
    index, sample = job
    fetch = synthetic_fetch(_REF_STATE["workload"], _REF_STATE["tables"], _REF_STATE["workload"].seed + 7919 * index)
    return run_reference_port(_REF_STATE["tables"], sample, fetch)


def reference_arm(args):
    """The reference's CPU implementation of the path on this box's host cores.

    The reference cannot be installed here (no qiskit wheels), so this is the oracle's literal Python port of its
    loop.  The reference itself is single-threaded; "all the host threads it can use" is realised the only way a
    user could: one process per core, each reconstructing its own PUBs (they are independent).
    """
    import multiprocessing as mp

    import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = W.get_workload(args.workload, args.scale)
    tables = W.build_tables(w)
    n_procs = max(1, int(os.environ.get("QCUT_REF_PROCS", "0") or 0) or (os.cpu_count() or 1))
    # a bounded sample per step and process so that warmup + steps end within a few minutes
    seconds = min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup))
    sample = cpu_sample(w, tables, None, seconds * n_procs, count=n_procs * max(1, -(-64 // n_procs)))  # equal shares
    jobs = [(i, sample[i::n_procs]) for i in range(n_procs) if sample[i::n_procs]]
    _REF_STATE.update(workload=w, tables=tables)
    work = secs = 0.0
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        for _ in range(args.warmup):
            pool.map(_reference_worker, [(i, part[:1]) for i, part in jobs])
        for _ in range(args.steps):
            t0 = time.perf_counter()
            parts = pool.map(_reference_worker, jobs)
            secs += time.perf_counter() - t0  # wall clock of the slowest process
            work += sum(wk for wk, _ in parts)
    value = work / secs
    sample_desc = (f"{len(sample)} PUBs x ~{sample[0][1]} shots of {args.workload} per step over {len(jobs)} processes "
                   f"(python port of the reference loop, one process per core)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8->int64 tallies, f64 sums", "data": "synthetic",
        "config": dict(w.describe(), l2="n/a (cpu)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(jobs), "kind": "port", "sample": sample_desc},
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


def ours_arm(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    import workloads as W
    from qiskit_addon_cutting_b200 import engine, reconstruct_expectation_values

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    entry.build_library()  # serialised across ranks by a file lock; a no-op when the library is current
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

    w = W.get_workload(args.workload, args.scale)
    tables = W.build_tables(w, rank, world)
    data = W.device_data(tables, w.seed * 1000 + rank)
    batch = engine.DeviceBatch(tables, data, flags=args.plan_flags)
    out = batch.out[: tables.n_obs]

    def step():
        n = batch.launch_tally()
        n += batch.launch_reduce()
        if world > 1:
            dist.all_reduce(out)
        return n

    for _ in range(max(args.warmup, 3)):
        launches = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(args.steps):
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
    elapsed = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device="cuda")
    k1 = torch.tensor([np.mean([e[0].elapsed_time(e[1]) for e in ev])], dtype=torch.float64, device="cuda")
    k2_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    if world > 1:
        dist.barrier()
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1, op=dist.ReduceOp.MAX)
    ms_total, k1_ms = float(elapsed.item()), float(k1.item())
    ms_per_step = ms_total / args.steps
    work_step = tables.work * world
    value = work_step / (ms_per_step * 1e-3)
    result_sample = out.cpu().numpy()

    # ---- roofline of the dominant kernel (K1): algorithmic bytes of one launch / its mean duration ----
    nbq = tables.pubs["nb_qpd"].astype(np.int64)
    nbo = tables.groups["nb_obs"][tables.pubs["group"]].astype(np.int64)
    k1_bytes = int((tables.pubs["shots"].astype(np.int64) * (nbo + nbq)).sum()) + 8 * tables.n_tallies
    peak, peak_src = peaks()
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    for path in sorted((ROOT / "profiles").glob("*_k1_traffic.json")):
        rec = json.loads(path.read_text())
        if rec.get("workload") == args.workload and abs(rec.get("scale", 1.0) - args.scale) < 1e-9:
            traffic = rec["traffic_bytes_per_k1_launch"]  # dram read + write of one qcut_tally_launch (ncu --set full)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "K1 parity tally (all launches of qcut_tally_launch)",
                "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": k1_ms, "peak_source": peak_src,
                "k1_share_of_step": k1_ms / ms_per_step, "k2_plus_allreduce_ms": k2_ms}

    # ---- end to end through the public API with host buffers -------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = np.empty(tables.data_bytes, dtype=np.uint8)
        torch.from_numpy(host).copy_(data[: tables.data_bytes])
        results = W.host_results(w, tables, host)
        coeffs_all = w.coefficients(rank, world)
        lo = rank * w.n_coeffs
        pairs = W.coefficient_pairs(coeffs_all[lo: lo + w.n_coeffs])
        got = reconstruct_expectation_values(results, pairs, w.observables)  # warm-up (pinned buffer, plan cache)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(args.e2e_steps):
            got = reconstruct_expectation_values(results, pairs, w.observables)
        secs = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        e2e_value = work_step * args.e2e_steps / float(secs.item())
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tables.data_bytes + tables.pubs.nbytes),
               "d2h_bytes_per_step": int(8 * tables.n_obs), "ms_per_step": 1e3 * float(secs.item()) / args.e2e_steps,
               "note": "each rank reconstructs its own shard through the public API (no process group)"}
        # same bytes, same path: the API result must equal the resident-data result of this rank's shard
        batch.run()
        torch.cuda.synchronize()
        assert np.array_equal(np.array(got), batch.out[: tables.n_obs].cpu().numpy()), "e2e result differs from resident result"
        del results, host

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port of the reference loop on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = cpu_sample(w, tables, None, args.cpu_seconds)

        def fetch(p, shots):
            pub = tables.pubs[p]
            nb = int(tables.groups[pub["group"]]["nb_obs"])
            o = data[int(pub["obs_offset"]): int(pub["obs_offset"]) + shots * nb].cpu().numpy().reshape(shots, nb)
            q = data[int(pub["qpd_offset"]): int(pub["qpd_offset"]) + shots * int(pub["nb_qpd"])].cpu().numpy().reshape(shots, -1)
            return o, q

        wk, secs = run_reference_port(tables, sample, fetch)
        compiled = None
        try:  # extra data point: the C restatement of the same loop (sequential fp64 included), all host threads
            from oracle import c_oracle

            n_sub = min(len(tables.pubs), 256)
            sub = tables.pubs[:n_sub].copy()
            hi = int(sub["qpd_offset"][-1]) + int(sub["shots"][-1]) * int(sub["nb_qpd"][-1]) + 16
            sub_bytes = data[:hi].cpu().numpy()
            sub["tally_offset"] -= sub["tally_offset"][0]
            n_t = int((tables.groups["n_terms"][sub["group"]]).sum())
            t0 = time.perf_counter()
            c_oracle.pubs(sub_bytes, sub, tables.groups, tables.masks, n_t, want_seq=True, n_threads=0)
            dt = time.perf_counter() - t0
            compiled = {"value": float((sub["shots"].astype(np.int64) * tables.groups["n_terms"][sub["group"]]).sum()) / dt,
                        "unit": UNIT, "cores": c_oracle.max_threads(), "kind": "port (oracle_seq.c, pthreads over PUBs)",
                        "sample": f"first {n_sub} PUBs of the resident data"}
        except Exception as exc:  # the C oracle is optional test infrastructure
            compiled = {"unavailable": repr(exc)}
        cpu = {"value": wk / secs, "unit": UNIT, "cores": 1, "kind": "port", "compiled_port": compiled,
               "sample": f"{len(sample)} PUBs x ~{sample[0][1]} shots of the same resident data "
                         f"({secs:.1f} s of the python port of the reference loop; the reference is single-threaded)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 shot rows -> int64 tallies -> f64 sums", "data": "synthetic",
            "config": dict(w.describe(), l2="inputs_exceed_l2" if tables.data_bytes > 2 * 126e6 else "inputs_fit_l2",
                           bytes_per_gpu=int(tables.data_bytes), work_per_gpu=int(tables.work)),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches) * args.steps,
            "gpu_launches_per_step": int(launches), "clocks": clocks,
            "checksum": float(np.sum(result_sample)),
        }
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
