"""Generate golden fixtures by executing the reference's OWN source text.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Qiskit is not installable here, so the reference package cannot be imported.
Instead this script

1. registers ``qiskit.quantum_info`` / ``qiskit.primitives`` shims whose
   ``Pauli`` / ``PauliList`` / result classes are the stand-ins from
   ``qiskit_addon_cutting_b200.containers``;
2. imports the reference's ``utils/observable_grouping.py`` unmodified from
   its file (it only depends on numpy and ``qiskit.quantum_info``);
3. ``exec``s, unmodified, the text of ``reconstruct_expectation_values``,
   ``_process_outcome``, ``_process_outcome_v2``, ``_outcome_to_int``
   (``cutting_reconstruction.py:31-232``), ``decompose_observables``
   (``cutting_decomposition.py:237-260``) and ``_get_pauli_indices``
   (``cutting_experiments.py:362-370``) in a namespace built from 1 and 2;
4. runs them on the reference's known-answer inputs and on the seeded cases of
   ``tests/_cases.py`` and writes inputs + outputs to ``tests/golden/*.json``.

Nothing from the reference is copied into the repository: only the numbers its
code produced.  Floats are stored with ``float.hex`` so fixtures are bit-exact.
"""

from __future__ import annotations

import base64
import json
import sys
from collections.abc import Mapping
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from qiskit_addon_cutting_b200 import containers as C  # noqa: E402
import _cases  # noqa: E402


from oracle.ref_loader import load_reference, load_reference_terms  # noqa: E402


def b64(a: np.ndarray) -> str:
    return base64.b64encode(np.ascontiguousarray(a).tobytes()).decode()


def dump_label(label):
    return list(label) if isinstance(label, tuple) else label


def dump_case(name: str, ref: dict, results, coefficients, observables) -> dict:
    partitioned = isinstance(observables, Mapping)
    obs_dict = observables if partitioned else {"A": observables}
    res_dict = results if partitioned else {"A": results}
    expvals = ref["reconstruct_expectation_values"](results, coefficients, observables)
    labels = []
    for label, obs in obs_dict.items():
        coll = ref["ObservableCollection"](obs)
        groups = [
            {
                "general": cog.general_observable.to_label(),
                "members": [p.to_label() for p in cog.commuting_observables],
                "pauli_indices": list(cog.pauli_indices),
                "pauli_bitmasks": [hex(m) for m in cog.pauli_bitmasks],
            }
            for cog in coll.groups
        ]
        pubs = []
        for pub in res_dict[label]:
            o, q = pub.data.observable_measurements, pub.data.qpd_measurements
            pubs.append(
                {
                    "shots": int(q.array.shape[0]),
                    "obs_bits": o.num_bits,
                    "qpd_bits": q.num_bits,
                    "obs": b64(o.array),
                    "qpd": b64(q.array),
                }
            )
        labels.append(
            {
                "label": dump_label(label),
                "observables": obs.to_labels(),
                "lookup": [[list(e) for e in coll.lookup[p]] for p in obs],
                "groups": groups,
                "pubs": pubs,
            }
        )
    return {
        "name": name,
        "partitioned": partitioned,
        "coefficients": [[float(c).hex(), w.name] for c, w in coefficients],
        "labels": labels,
        "reference_expvals": [float(v).hex() for v in expvals],
    }


def dump_observable(o) -> dict:
    if isinstance(o, C.Pauli):
        return {"pauli": o.to_label()}
    return {"labels": o.paulis.to_labels(), "coeffs": [[float(c.real).hex(), float(c.imag).hex()] for c in o.coeffs]}


def main_terms() -> None:
    """Golden vectors for gather_unique_observable_terms / reconstruct_observable_expvals_from_terms."""
    mod = load_reference_terms()
    # the reference's own known answers (test/utils/test_observable_terms.py:61-91)
    evs = mod.reconstruct_observable_expvals_from_terms(
        [C.Pauli("XX"), C.SparsePauliOp(C.PauliList(["iXY", "ZZ", "XX"]))],
        {C.Pauli("XX"): 7, C.Pauli("XY"): 11, C.Pauli("ZZ"): 13},
    )
    assert evs == [7, 20 + 11j], evs
    assert mod.reconstruct_observable_expvals_from_terms([C.SparsePauliOp(["XX", "-XY"], [0, 1])], {C.Pauli("XY"): 3}) == [-3]
    assert mod.reconstruct_observable_expvals_from_terms(C.PauliList(["iXY", "-ZZ"]), {C.Pauli("XY"): 7, C.Pauli("ZZ"): 11}) == [7j, -11]
    assert mod.gather_unique_observable_terms([C.Pauli("-XX"), C.SparsePauliOp(C.PauliList(["iXY", "ZZ", "XX"]))]) == C.PauliList(["XX", "XY", "ZZ"])
    cases = {}
    for name, make in _cases.TERMS_CASES.items():
        observables, values = make()
        out = mod.reconstruct_observable_expvals_from_terms(observables, values)
        unique = mod.gather_unique_observable_terms(observables)
        cases[name] = {
            "observables": [dump_observable(o) for o in observables],
            "term_expvals": [[p.to_label(), float(np.real(v)).hex(), float(np.imag(v)).hex() if isinstance(v, complex) else None]
                             for p, v in values.items()],
            "unique_terms": unique.to_labels(),
            "reference_expvals": [[float(np.real(v)).hex(), float(np.imag(v)).hex()] for v in out],
        }
        print(f"terms/{name:12s} expvals[:2] = {[complex(v) for v in out[:2]]}")
    (HERE / "terms_cases.json").write_text(json.dumps(cases) + "\n")


def main() -> None:
    ref = load_reference()
    recon = ref["reconstruct_expectation_values"]

    # ---- the reference's own known-answer tests (test/test_cutting_reconstruction.py) ----
    kat = {}
    # :117-138  SamplerV2 golden vector
    obs_data = C.BitArray(np.array([0, 0, 0, 1, 0, 0, 2, 3, 1, 0], dtype=np.uint8).reshape(-1, 1), 2)
    qpd_data = C.BitArray(np.array([0, 1, 1, 3, 2, 0, 1, 3, 0, 2], dtype=np.uint8).reshape(-1, 1), 2)
    results = C.PrimitiveResult(
        [C.PubResult(C.DataBin(observable_measurements=obs_data, qpd_measurements=qpd_data))]
    )
    out = recon(results, [(1.0, C.WeightType.EXACT)], C.PauliList(["II", "IZ", "ZI", "ZZ"]))
    assert np.allclose(out, [0.0, -0.6, 0.0, -0.2], rtol=1e-6, atol=1e-12), out
    kat["v2_golden"] = {"expected_approx": [0.0, -0.6, 0.0, -0.2], "reference_expvals": [float(v).hex() for v in out]}
    # :43-50  SamplerV1 trivial
    v1 = C.SamplerResult(quasi_dists=[C.QuasiDistribution({"0": 1.0})], metadata=[{}])
    out = recon(v1, [(1.0, C.WeightType.EXACT)], C.PauliList(["ZZ"]))
    assert out == [1.0]
    kat["v1_trivial"] = {"reference_expvals": [float(v).hex() for v in out]}
    # :152-170  _process_outcome truth table
    cog = ref["CommutingObservableGroup"](C.Pauli("XZ"), list(C.PauliList(["IZ", "XI", "XZ"])))
    table = {}
    expected = {
        "000": [1, 1, 1], "001": [-1, 1, -1], "010": [1, -1, -1], "011": [-1, -1, 1],
        "100": [-1, -1, -1], "101": [1, -1, 1], "110": [-1, 1, 1], "111": [1, 1, -1],
    }  # fmt: skip
    for outcome, exp in expected.items():
        for o in (outcome, f"0b{outcome}", int(f"0b{outcome}", 0), hex(int(f"0b{outcome}", 0))):
            got = ref["_process_outcome"](cog, o)
            assert np.all(got == exp), (o, got, exp)
        table[outcome] = [int(v) for v in got]
    kat["process_outcome"] = {"general": "XZ", "members": ["IZ", "XI", "XZ"], "masks": list(cog.pauli_bitmasks), "table": table}
    # :139-150  count-mismatch message
    try:
        recon(v1, [(1.0, C.WeightType.EXACT)], C.PauliList(["ZZ", "XX"]))
        raise AssertionError("expected ValueError")
    except ValueError as e:
        kat["count_mismatch_message"] = e.args[0]
    # test/utils/test_observable_grouping.py:104-107  group sizes
    coll = ref["ObservableCollection"](C.PauliList(["XX", "ZZ", "XX"]))
    kat["group_sizes_XX_ZZ_XX"] = [len(g.commuting_observables) for g in coll.groups]
    (HERE / "reference_kat.json").write_text(json.dumps(kat, indent=1) + "\n")

    # ---- V1 case with non-trivial quasi-distributions ----
    rng = np.random.default_rng(11)
    observables = {"A": C.PauliList(["ZI", "IZ", "XX"]), "B": C.PauliList(["ZZ", "II", "XI"])}
    coeffs = _cases.random_coefficients(rng, 3)
    v1_results = {}
    from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection as OurCollection

    for label, obs in observables.items():
        groups = OurCollection(obs).groups
        dists = []
        for _ in range(3 * len(groups)):
            keys = rng.choice(1 << 4, size=6, replace=False)
            probs = rng.normal(size=6)
            dists.append(C.QuasiDistribution({int(k): float(p) for k, p in zip(keys, probs)}))
        v1_results[label] = C.SamplerResult(dists)
    out = recon(v1_results, coeffs, observables)
    v1_case = {
        "coefficients": [[float(c).hex(), w.name] for c, w in coeffs],
        "labels": [
            {
                "label": label,
                "observables": obs.to_labels(),
                "quasi_dists": [[[int(k), float(p).hex()] for k, p in d.items()] for d in v1_results[label].quasi_dists],
            }
            for label, obs in observables.items()
        ],
        "reference_expvals": [float(v).hex() for v in out],
    }
    (HERE / "v1_case.json").write_text(json.dumps(v1_case, indent=1) + "\n")

    # ---- seeded V2 cases ----
    for name, make in _cases.SMALL_CASES.items():
        results, coefficients, observables = make()
        case = dump_case(name, ref, results, coefficients, observables)
        (HERE / f"case_{name}.json").write_text(json.dumps(case) + "\n")
        print(f"{name:14s} expvals[:3] = {[float.fromhex(v) for v in case['reference_expvals'][:3]]}")


if __name__ == "__main__":
    if sys.argv[1:] == ["terms"]:
        load_reference()
        main_terms()
    else:
        main()
        main_terms()
