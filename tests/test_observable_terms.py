"""Term recombination (reference utils/observable_terms.py, test/utils/test_observable_terms.py).

CPU: the oracle restatement and the host-side gather reproduce the fixtures that the reference's own source
produced (tests/golden/make_golden.py terms), and the reference's known answers.  GPU: ``qcut_recombine_launch``
behind ``reconstruct_observable_expvals_from_terms`` equals them bit for bit.
"""

import json
from pathlib import Path

import numpy as np
import pytest

import _cases
from oracle import oracle
from qiskit_addon_cutting_b200 import (
    Pauli,
    PauliList,
    SparsePauliOp,
    TermRecombination,
    gather_unique_observable_terms,
    reconstruct_observable_expvals_from_terms,
)

GOLDEN = json.loads((Path(__file__).parent / "golden" / "terms_cases.json").read_text())


def _load(name):
    case = GOLDEN[name]
    observables = [
        Pauli(o["pauli"]) if "pauli" in o
        else SparsePauliOp(o["labels"], [complex(float.fromhex(r), float.fromhex(i)) for r, i in o["coeffs"]])
        for o in case["observables"]
    ]
    values = {
        Pauli(lab): (float.fromhex(re) if im is None else complex(float.fromhex(re), float.fromhex(im)))
        for lab, re, im in case["term_expvals"]
    }
    want = [complex(float.fromhex(r), float.fromhex(i)) for r, i in case["reference_expvals"]]
    return observables, values, want, case["unique_terms"]


def _bits(values) -> bytes:
    return np.array([complex(v) for v in values], dtype=np.complex128).tobytes()


KATS = [  # test/utils/test_observable_terms.py:61-91
    (lambda: [Pauli("XX"), SparsePauliOp(PauliList(["iXY", "ZZ", "XX"]))], {"XX": 7, "XY": 11, "ZZ": 13}, [7, 20 + 11j]),
    (lambda: [SparsePauliOp(["XX", "-XY"], [0, 1])], {"XY": 3}, [-3]),
    (lambda: PauliList(["iXY", "-ZZ"]), {"XY": 7, "ZZ": 11}, [7j, -11]),
]


@pytest.mark.parametrize("name", list(GOLDEN))
def test_seeded_cases_are_the_committed_ones(name):
    """The fixtures were generated from exactly the inputs tests/_cases.py builds today."""
    observables, values, _, _ = _load(name)
    made_obs, made_values = _cases.TERMS_CASES[name]()
    assert len(made_obs) == len(observables)
    for a, b in zip(made_obs, observables):
        assert type(a) is type(b)
        if isinstance(a, Pauli):
            assert a == b
        else:
            assert a.paulis == b.paulis and np.array_equal(a.coeffs, b.coeffs)
    assert {p.to_label(): v for p, v in made_values.items()} == {p.to_label(): v for p, v in values.items()}


@pytest.mark.parametrize("name", list(GOLDEN))
def test_oracle_matches_reference_bit_for_bit(name):
    observables, values, want, unique = _load(name)
    assert _bits(oracle.observable_expvals_from_terms(observables, values)) == _bits(want)
    labels = [Pauli((np.array(z), np.array(x))).to_label() for z, x in oracle.unique_term_keys(observables)]
    assert labels == unique


def test_oracle_known_answers():
    for make, terms, want in KATS:
        assert oracle.observable_expvals_from_terms(make(), {Pauli(k): v for k, v in terms.items()}) == want
    with pytest.raises(ValueError) as e_info:
        oracle.observable_expvals_from_terms([Pauli("XY")], {Pauli("XX"): 5, Pauli("ZZ"): 7})
    assert e_info.value.args[0] == "An observable contains a Pauli term whose expectation value was not provided."


@pytest.mark.parametrize("name", list(GOLDEN))
def test_gather_matches_reference(name):
    observables, _, _, unique = _load(name)
    got = gather_unique_observable_terms(observables)
    assert got.to_labels() == unique and not np.any(got.phase)


def test_gather_known_answers_and_errors():
    # test/utils/test_observable_terms.py:25-59
    terms = gather_unique_observable_terms([Pauli("-XX"), SparsePauliOp(PauliList(["iXY", "ZZ", "XX"]))])
    assert terms.num_qubits == 2 and terms == PauliList(["XX", "XY", "ZZ"])
    assert gather_unique_observable_terms(SparsePauliOp(PauliList(["XYZ"]), [0])) == PauliList(["III"])[:0]
    assert gather_unique_observable_terms(PauliList(["XXXX"])[:0]) == PauliList(["IIII"])[:0]
    with pytest.raises(ValueError) as e_info:
        gather_unique_observable_terms([Pauli("XX"), Pauli("XYZ")])
    assert e_info.value.args[0] == "Cannot construct PauliList.  Do provided observables all have the same number of qubits?"
    with pytest.raises(ValueError) as e_info:
        gather_unique_observable_terms([])
    assert e_info.value.args[0] == "observables list cannot be empty"


def test_rows_skip_zero_coefficients_and_absorb_phases():
    r = TermRecombination([Pauli("-XX"), SparsePauliOp(["XX", "-XY", "iZZ"], [0, 1, 2])],
                          {Pauli("XX"): 0, Pauli("XY"): 1, Pauli("ZZ"): 2})
    assert r.row_start.tolist() == [0, 1, 3] and r.index.tolist() == [0, 1, 2]
    assert r.coeffs.tolist() == [-1, -1, 2j]


def test_host_side_errors_need_no_device():
    assert reconstruct_observable_expvals_from_terms([], {Pauli("XX"): 7}) == []
    assert reconstruct_observable_expvals_from_terms([], {}) == []
    with pytest.raises(ValueError) as e_info:
        reconstruct_observable_expvals_from_terms([Pauli("XY")], {Pauli("XX"): 5, Pauli("ZZ"): 7})
    assert e_info.value.args[0] == "An observable contains a Pauli term whose expectation value was not provided."


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from qiskit_addon_cutting_b200._lib import QcutError

    with pytest.raises(QcutError):
        reconstruct_observable_expvals_from_terms([Pauli("XX")], {Pauli("XX"): 5})


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GOLDEN))
def test_gpu_matches_reference_bit_for_bit(name):
    observables, values, want, _ = _load(name)
    got = reconstruct_observable_expvals_from_terms(observables, values)
    assert _bits(got) == _bits(want)
    assert all(isinstance(v, np.complex128) for v in got)


@pytest.mark.gpu
def test_gpu_known_answers():
    for make, terms, want in KATS:
        assert reconstruct_observable_expvals_from_terms(make(), {Pauli(k): v for k, v in terms.items()}) == want


@pytest.mark.gpu
@pytest.mark.parametrize("complex_values", [False, True])
def test_gpu_recombination_of_resident_term_values(complex_values):
    """Values that are already in HBM (the output vector of the reconstruction) and long rows."""
    import torch

    rng = np.random.default_rng(77)
    paulis = list(_cases.random_paulis(rng, 300, 12, 1, 6).unique())
    index = {Pauli((p.z, p.x)): i for i, p in enumerate(paulis)}
    observables = []
    for _ in range(40):
        n = int(rng.integers(1, 2500))
        pick = rng.integers(len(paulis), size=n)
        observables.append(SparsePauliOp(PauliList([paulis[i] for i in pick]), rng.normal(size=n) + 1j * rng.normal(size=n)))
    values = rng.uniform(-1, 1, size=len(paulis)) + (1j * rng.uniform(-1, 1, size=len(paulis)) if complex_values else 0)
    values = values if complex_values else values.real.copy()
    got = TermRecombination(observables, index).launch(torch.from_numpy(values).cuda()).cpu().numpy()
    want = oracle.observable_expvals_from_terms(observables, {p: values[i] for p, i in index.items()})
    assert _bits(got) == _bits(want)
