"""Markdown tables (per-config throughput / roofline and parity deviations) from a bench.py JSON line.

    python scripts/bench_tables.py profiles/r02_bench_default.json
"""
import json
import sys

line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
recs = [line] + [r for r in line.get("per_config", []) if "error" not in r]

print("| workload | shot bytes | K1 ms | GB/s | of peak | step ms | shot·terms/s | e2e ms/call | e2e shot·terms/s | reference CPU (1 core) | reference-rounding mode ms |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in recs:
    rf, e, c, lit = r["roofline"], r.get("e2e") or {}, r.get("cpu_baseline") or {}, r.get("reference_rounding_mode") or {}
    print(f"| {r['config']['workload']} | {r['config']['bytes_per_gpu'] / 1e9:.3g} GB | {rf['kernel_ms']:.4g} | {rf['achieved']:.0f} | "
          f"{rf['frac']:.3f} | {r['ms_per_step']:.4g} | {r['value']:.3g} | {e.get('ms_per_step', float('nan')):.4g} | "
          f"{e.get('value', float('nan')):.3g} | {c.get('value', float('nan')):.3g} | {lit.get('ms_per_step', float('nan')):.4g} |")
print()
print("| workload | PUBs checked (tallies bit-exact) | sample (tuples x shots) | max \\|E\\| per term | ours − reference | ours − exact | reference − exact | "
      "reference arithmetic − tally/S per term, full shots | reference-rounding mode == reference |")
print("|---|---|---|---:|---:|---:|---:|---:|---|")
for r in recs:
    p = r["parity"]
    sample = p.get("expvals_sample", "").split(" of ")[0].replace("first ", "")
    print(f"| {r['config']['workload']} | {p['pubs_checked']} ({p['tallies_bit_exact']}) | {sample} | {p.get('per_term_max_abs_expval', float('nan')):.3g} | "
          f"{p.get('max_abs_dev_vs_reference', float('nan')):.2e} | {p.get('max_abs_dev_vs_exact', float('nan')):.2e} | "
          f"{p.get('reference_vs_exact', float('nan')):.2e} | {p.get('per_term_max_abs_dev_reference_arithmetic_vs_exact', float('nan')):.2e} | "
          f"{p.get('reference_rounding_mode_bit_identical_to_reference')} |")
