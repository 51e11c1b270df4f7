"""Loader for the fixtures written by tests/golden/make_golden.py."""

from __future__ import annotations

import base64
import json
from pathlib import Path

import numpy as np

from qiskit_addon_cutting_b200.containers import (
    BitArray,
    DataBin,
    PauliList,
    PrimitiveResult,
    PubResult,
    QuasiDistribution,
    SamplerResult,
    WeightType,
)

GOLDEN = Path(__file__).resolve().parent / "golden"
CASE_NAMES = sorted(p.stem[len("case_"):] for p in GOLDEN.glob("case_*.json"))


def _label(raw):
    return tuple(raw) if isinstance(raw, list) else raw


def _array(b64: str, shots: int, bits: int) -> np.ndarray:
    nb = (bits + 7) // 8
    return np.frombuffer(base64.b64decode(b64), dtype=np.uint8).reshape(shots, nb).copy()


def load_kat() -> dict:
    return json.loads((GOLDEN / "reference_kat.json").read_text())


def load_case(name: str) -> dict:
    """Returns dict(results, coefficients, observables, reference_expvals, raw)."""
    raw = json.loads((GOLDEN / f"case_{name}.json").read_text())
    observables, results = {}, {}
    for lab in raw["labels"]:
        label = _label(lab["label"])
        observables[label] = PauliList(lab["observables"])
        pubs = []
        for pub in lab["pubs"]:
            pubs.append(
                PubResult(
                    DataBin(
                        observable_measurements=BitArray(_array(pub["obs"], pub["shots"], pub["obs_bits"]), pub["obs_bits"]),
                        qpd_measurements=BitArray(_array(pub["qpd"], pub["shots"], pub["qpd_bits"]), pub["qpd_bits"]),
                    )
                )
            )
        results[label] = PrimitiveResult(pubs)
    coefficients = [(float.fromhex(c), WeightType[w]) for c, w in raw["coefficients"]]
    if not raw["partitioned"]:
        observables, results = observables["A"], results["A"]
    return {
        "results": results,
        "coefficients": coefficients,
        "observables": observables,
        "reference_expvals": [float.fromhex(v) for v in raw["reference_expvals"]],
        "raw": raw,
    }


def load_v1_case() -> dict:
    raw = json.loads((GOLDEN / "v1_case.json").read_text())
    observables, results = {}, {}
    for lab in raw["labels"]:
        observables[lab["label"]] = PauliList(lab["observables"])
        results[lab["label"]] = SamplerResult(
            [QuasiDistribution({int(k): float.fromhex(p) for k, p in dist}) for dist in lab["quasi_dists"]]
        )
    return {
        "results": results,
        "coefficients": [(float.fromhex(c), WeightType[w]) for c, w in raw["coefficients"]],
        "observables": observables,
        "reference_expvals": [float.fromhex(v) for v in raw["reference_expvals"]],
    }
