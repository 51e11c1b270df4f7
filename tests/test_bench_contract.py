"""The committed bench lines (profiles/r01_bench_*.json, produced by bench.py on a B200) carry every key of the
measurement contract, so a change to bench.py that drops one is caught here without a GPU."""

import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parents[1] / "profiles"
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}


def _line(name):
    return json.loads((PROFILES / name).read_text().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r01_bench_default.json", "r01_bench_default_20steps_final.json"])
def test_default_bench_line(name):
    d = _line(name)
    assert BASE_KEYS <= set(d)
    assert d["metric"] == "reconstruct_expectation_values shots*terms/sec" and d["unit"] == "shot-terms/s"
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["config"]["workload"] == "cfg4_heavy_hex" and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["config"]["l2"] == "inputs_exceed_l2"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] >= 0.98 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] >= d["config"]["bytes_per_gpu"] and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]  # the end-to-end number is not the resident number repeated
    assert d["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert abs(d["value"] - d["config"]["work_per_gpu"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_reference_arm_line():
    d = _line("r01_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "reconstruct_expectation_values shots*terms/sec"
    assert d["config"]["workload"] == "cfg4_heavy_hex" and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
