"""Regenerate the round-2 tables at the end of DESIGN.md from the committed profiles (between the R02 markers).

    python scripts/design_tables.py
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PROF = ROOT / "profiles"


def line(path):
    return json.loads(Path(path).read_text().strip().splitlines()[-1])


def scaling_table():
    legs = (("strong_cfg4", "cfg4 strong (97 000 tuples x 8192 shots in total, 14.3 GB)", "strong"),
            ("strong_cfg4_s65536", "cfg4 strong, 65 536 shots per subexperiment (115 GB in total)", None),
            ("strong_cfg5", "cfg5 strong (8.2 GB in total)", "cfg5_strong"),
            ("weak_cfg5", "cfg5 weak (8.2 GB per GPU)", "cfg5_weak"))
    out = ["| leg | N | shot·terms/s | ms per step | K1 ms | K1 of HBM peak (per GPU) | K2 + all-reduce ms | efficiency vs N=1 | e2e shot·terms/s (sharded public API) | e2e ms per call | summed staging rate / host copy ceiling |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for leg, label, n2 in legs:
        base = None
        for n in (1, 2, 4, 8):
            path = PROF / (f"r02_n2_{n2}.json" if n == 2 else f"r02_scale_n{n}_{leg}.json")
            if n == 2 and n2 is None:
                continue
            try:
                d = line(path)
            except (OSError, ValueError, IndexError):
                continue
            if n == 1:
                base = d["value"]
            if base is None and leg == "weak_cfg5":
                base = line(PROF / "r02_scale_n1_strong_cfg5.json")["value"]
            eff = d["value"] / (base * n) if base else float("nan")
            e = d.get("e2e") or {}
            out.append(f"| {label} | {n} | {d['value']:.3g} | {d['ms_per_step']:.4g} | {d['roofline']['kernel_ms']:.4g} | {d['roofline']['frac']:.3f} | "
                       f"{d['roofline']['k2_plus_allreduce_ms']:.3f} | {eff:.2f} | {('%.3g' % e['value']) if e.get('value') else '—'} | "
                       f"{('%.4g' % e['ms_per_step']) if e.get('ms_per_step') else '—'} | {('%.2f' % e['host_copy_frac']) if e.get('host_copy_frac') else '—'} |")
    return "\n".join(out)


def cold_table(d):
    out = ["| workload | first public call of the process, empty kernel cache (ms) | NVRTC state when it returned | specialised kernels ready after (s) | steady end-to-end call (ms) |",
           "|---|---:|---|---:|---:|"]
    names = {0: "none needed (launch-bound call, ahead-of-time kernels)", 1: "still compiling in the background", 2: "ready"}
    for r in [d] + [r for r in d.get("per_config", []) if "error" not in r]:
        c = r.get("cold_start") or {}
        if "first_call_ms" not in c:
            continue
        ready = c.get("specialised_kernels_ready_after_s")
        out.append(f"| {r['config']['workload']} | {c['first_call_ms']:.1f} | {names.get(c.get('jit_state_after_first_call'), '?')} | "
                   f"{('%.2f' % ready) if ready else '—'} | {(r.get('e2e') or {}).get('ms_per_step', float('nan')):.4g} |")
    return "\n".join(out)


def ncu_table():
    import re

    names = [("k1_cfg4", "K1 specialised, cfg4 (8+1 B rows, 134 terms)"), ("k1_cfg5", "K1 cfg5: its 4 switch kernels, staged stream (scale 0.5)"),
             ("k1_cfg3_local", "K1 cfg3 local (2 mixed-width kernels, 1-7 B rows)"), ("k1_cfg3_dense", "K1 cfg3 dense (100 terms x 32 planes)"),
             ("k1_wide_slab", "K1 16-byte rows in 8-byte slabs (271 terms)"), ("k1_wide", "K1 wide rows, shared-memory engine (16 B, 271 terms)"), ("k1_rt", "K1 run-time masks (ahead of time), cfg4 rows"),
             ("k1_generic", "K1 generic (AND/POPC per shot), cfg4 rows"), ("k2_stream", "K2 stream (cfg4, 97 000 tuples)"),
             ("literal", "K1-literal (reference rounding), cfg4 rows")]
    out = ["| kernel (`profiles/r02_<name>_ncu_summary.txt`) | launches | time | DRAM % of peak | issue slots % | ALU / XU / FMA pipe % | registers | warps per scheduler | global-load sectors per request | requested / ideal L2 sectors | L2->SM sectors per DRAM sector | top stalls |",
           "|---|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---|"]
    for name, label in names:
        path = PROF / f"r02_{name}_ncu_summary.txt"
        if not path.exists():
            continue
        blocks = path.read_text().split("=== kernel")[1:]

        def val(block, key):
            m = re.search(re.escape(key) + r"\s+\S+\s+([\d.,]+)", block)
            return float(m.group(1).replace(",", "")) if m else float("nan")

        def unit(block, key):
            m = re.search(re.escape(key) + r"\s+(\S+)", block)
            return m.group(1) if m else ""

        times = [val(b, "gpu__time_duration.sum") * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(unit(b, "gpu__time_duration.sum"), 1.0) for b in blocks]
        w = [t / sum(times) for t in times]

        def avg(key):
            return sum(val(b, key) * wi for b, wi in zip(blocks, w))

        def derived(i):
            vals = []
            for b, wi in zip(blocks, w):
                m = re.search(r"sectors per request ([\d.]+); L2->SM read sectors / DRAM read sectors ([\d.]+); requested / ideal L2 sectors of global accesses ([\d.]+)", b)
                vals.append(float(m.group(i)) * wi if m else float("nan"))
            return sum(vals)

        stalls = re.search(r"stalls \(warps per issue-active cycle\): (.*)", blocks[max(range(len(blocks)), key=lambda i: times[i])])
        top = ", ".join(stalls.group(1).split(", ")[:3]) if stalls else ""
        out.append(f"| {label} | {len(blocks)} | {sum(times):.3f} ms | {avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                   f"{avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {avg('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.0f} / "
                   f"{avg('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.0f} / {avg('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.0f} | "
                   f"{avg('launch__registers_per_thread'):.0f} | {avg('smsp__warps_active.avg.per_cycle_active'):.2f} | {derived(1):.1f} | {derived(3):.3f} | {derived(2):.2f} | {top} |")
    return "\n".join(out)


def main():
    bench = PROF / "r02_bench_default.json"
    d = line(bench)
    tables = subprocess.run([sys.executable, str(ROOT / "scripts" / "bench_tables.py"), str(bench)], capture_output=True, text=True, check=True).stdout
    dd = d.get("dedup") or {}
    text = [f"### Per shape, one GPU (`profiles/{bench.name}`: the driver's command, fresh kernel cache, biased synthetic data)\n", tables,
            "### Cold start (same run)\n", cold_table(d), ""]
    if "dedup_ratio" in dd:
        text += ["### PUB de-duplication (same run)\n",
                 f"{dd['workload']}: {dd['pubs']} PUBs handed over, {dd['with_dedup']['pubs_tallied']} staged and tallied; H2D shot bytes "
                 f"{dd['without_dedup']['h2d_shot_bytes']:,} -> {dd['with_dedup']['h2d_shot_bytes']:,} (`dedup_ratio` {dd['dedup_ratio']:.1f}); "
                 f"{dd['without_dedup']['ms_per_call']:.2f} -> {dd['with_dedup']['ms_per_call']:.2f} ms per call; results bit-identical: {dd['results_bit_identical']}.\n"]
    text += ["### ncu `--set full --clock-control none` per kernel (time-weighted over the captured launches; cold-cache, serialised)\n",
             ncu_table(), "",
             "Reading the sector columns: the shot-byte stream is read with warp-wide 128-bit loads (16 sectors per request); the average "
             "per request is lower because descriptor, qpd (2-byte) and mask loads count too.  `requested / ideal` = 1.01 on the 8-byte "
             "shapes: no partially used sectors.  The 1-7-byte shapes of cfg3 local / cfg5 read 16 consecutive shots per lane, so "
             "neighbouring lanes' 16-byte pieces share 32-byte sectors (1.2-1.9 requested / ideal at L1); those loads go through L1 "
             "(allocating) and L2->SM traffic stays within 1.1-1.5x of DRAM traffic.  DRAM bytes per launch on cfg4: 14.73 GB vs 14.51 GB "
             "algorithmic (`profiles/r02_k1_traffic.json`).\n"]
    text += ["### Scaling on one box (`profiles/r02_scale_n*`, `profiles/r02_n2_*`; the driver's own SCALE run covers weak cfg4)\n", scaling_table(), ""]
    design = (ROOT / "DESIGN.md").read_text()
    begin, end = "<!-- R02-TABLES-BEGIN -->", "<!-- R02-TABLES-END -->"
    block = begin + "\n" + "\n".join(text) + "\n" + end
    if begin in design:
        design = design[: design.index(begin)] + block + design[design.index(end) + len(end):]
    else:
        design = design.rstrip("\n") + "\n\n## 9. Round-2 tables (generated by `scripts/design_tables.py` from `profiles/`)\n\n" + block + "\n"
    (ROOT / "DESIGN.md").write_text(design)
    print("DESIGN.md tables regenerated from", bench.name)


if __name__ == "__main__":
    main()
