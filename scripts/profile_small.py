"""Where does a small reconstruction call spend its host time?  cProfile over many calls (GPU box).

    python scripts/profile_small.py cfg1_tutorial 300
"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import torch  # noqa: E402

import workloads as W  # noqa: E402
from qiskit_addon_cutting_b200 import reconstruct_expectation_values  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_tutorial"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 300
w = W.get_workload(name, 1.0)
tables = W.build_tables(w)
data = W.device_data(tables, 1)
results = W.host_results(w, tables, W.device_fetch(data))
pairs = W.coefficient_pairs(tables.coeffs)
for _ in range(5):
    reconstruct_expectation_values(results, pairs, w.observables)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(calls):
    reconstruct_expectation_values(results, pairs, w.observables)
print(f"{name}: {1e3 * (time.perf_counter() - t) / calls:.3f} ms per call")
prof = cProfile.Profile()
prof.enable()
for _ in range(calls):
    reconstruct_expectation_values(results, pairs, w.observables)
prof.disable()
stats = pstats.Stats(prof)
stats.sort_stats("tottime").print_stats(28)
