"""bench.py's measurement contract, checked by RUNNING what can run without a GPU:

* the reference arm (``bench.py --impl reference``): executes the reference's own source text on a bounded sample with
  one process per core and must print ONE JSON line with the contract's keys;
* the sample sizing and the sharding helpers both arms share (pure host code);
* the per-config / parity tables are rendered from a line by ``scripts/bench_tables.py``.

The GPU arm itself needs a device; the driver runs it at round end.  Its latest committed line (``profiles/``) is only
used here as an input for the table renderer."""

import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import workloads as W  # noqa: E402


def test_reference_arm_runs_and_prints_one_contract_line():
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--workload", "cfg2_wire", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=300,
                          env={**__import__("os").environ, "QCUT_REF_PROCS": "2"})
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.strip().splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["config"]["workload"] == "cfg2_wire" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["value"] == d["value"] and c["cores"] == 2 and c["kind"] in ("reference", "port") and c["sample"]
    from oracle import ref_loader

    assert c["kind"] == ("reference" if ref_loader.available() else "port")


def test_sample_sizing_is_bounded_and_within_the_workload():
    for name in W.WORKLOADS:
        w = W.get_workload(name, 0.05 if name.startswith("cfg4") else 1.0)
        for seconds in (0.5, 4.0, 12.0):
            c, s = bench.size_sample(w, seconds)
            assert 1 <= c <= w.n_coeffs and 1 <= s <= w.shots
            per_shot = sum(bench.REF_SHOT_US + bench.REF_TERM_US * len(g.pauli_bitmasks) for col in w.collections() for g in col.groups)
            assert c * s * per_shot * 1e-6 <= max(2.5 * seconds, 2 * 64 * per_shot * 1e-6)
            assert bench.sample_work(w, c, s) == c * s * sum(len(g.pauli_bitmasks) for col in w.collections() for g in col.groups)


@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_shards_partition_the_global_problem(scaling):
    w = W.get_workload("cfg5_molecular", 1.0)
    world = 8
    seen, total = [], None
    for rank in range(world):
        local, coeffs, (lo, hi) = W.shard(w, rank, world, scaling)
        total = len(coeffs)
        assert local.n_coeffs == hi - lo
        seen.append((lo, hi))
        tables = W.build_tables(local, coeffs=coeffs[lo:hi])
        assert len(tables.coeffs) == hi - lo and np.array_equal(tables.coeffs, coeffs[lo:hi])
    assert seen[0][0] == 0 and seen[-1][1] == total and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    assert total == (w.n_coeffs * world if scaling == "weak" else w.n_coeffs)


def test_tables_render_from_a_committed_gpu_line():
    lines = sorted((ROOT / "profiles").glob("r02_bench_default*.json"))
    if not lines:
        pytest.skip("no round-2 bench line committed yet")
    proc = subprocess.run([sys.executable, str(ROOT / "scripts" / "bench_tables.py"), str(lines[-1])], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "cfg4_heavy_hex" in proc.stdout and "cfg5_molecular" in proc.stdout and "reference-rounding" in proc.stdout
