"""The oracle is pinned against the reference's own KATs and against fixtures its own source produced."""

import numpy as np
import pytest

import _golden
from oracle import c_oracle, oracle
from qiskit_addon_cutting_b200.containers import (
    BitArray, DataBin, Pauli, PauliList, PrimitiveResult, PubResult, QuasiDistribution, SamplerResult, WeightType,
)
from qiskit_addon_cutting_b200.observable_grouping import CommutingObservableGroup, ObservableCollection

KAT = _golden.load_kat()


def _kat_v2_inputs():
    obs = BitArray(np.array([0, 0, 0, 1, 0, 0, 2, 3, 1, 0], dtype=np.uint8).reshape(-1, 1), 2)
    qpd = BitArray(np.array([0, 1, 1, 3, 2, 0, 1, 3, 0, 2], dtype=np.uint8).reshape(-1, 1), 2)
    results = PrimitiveResult([PubResult(DataBin(observable_measurements=obs, qpd_measurements=qpd))])
    return results, [(1.0, WeightType.EXACT)], PauliList(["II", "IZ", "ZI", "ZZ"])


def test_reference_v2_golden_vector():
    """test/test_cutting_reconstruction.py:117-138 -> approx [0.0, -0.6, 0.0, -0.2]."""
    out = oracle.reconstruct_literal(*_kat_v2_inputs())
    assert out == pytest.approx([0.0, -0.6, 0.0, -0.2])
    # and bit-for-bit what the reference's own source produced here
    assert [float(v).hex() for v in out] == KAT["v2_golden"]["reference_expvals"]
    assert oracle.reconstruct_exact(*_kat_v2_inputs()) == pytest.approx([0.0, -0.6, 0.0, -0.2], abs=1e-15)


def test_reference_v1_trivial():
    """test/test_cutting_reconstruction.py:43-50."""
    results = SamplerResult(quasi_dists=[QuasiDistribution({"0": 1.0})], metadata=[{}])
    assert oracle.reconstruct_literal(results, [(1.0, WeightType.EXACT)], PauliList(["ZZ"])) == [1.0]


def test_reference_process_outcome_truth_table():
    """test/test_cutting_reconstruction.py:152-170: cog XZ with members IZ, XI, XZ -> masks 1, 2, 3."""
    pauli_indices, masks = oracle.group_masks(list(PauliList(["IZ", "XI", "XZ"])))
    assert masks == KAT["process_outcome"]["masks"] == [1, 2, 3]
    assert pauli_indices == [0, 1]
    for outcome, expected in KAT["process_outcome"]["table"].items():
        for o in (outcome, f"0b{outcome}", int(f"0b{outcome}", 0), hex(int(f"0b{outcome}", 0))):
            assert np.all(oracle.process_outcome(masks, len(pauli_indices), o) == expected)
    cog = CommutingObservableGroup(Pauli("XZ"), list(PauliList(["IZ", "XI", "XZ"])))
    assert cog.pauli_bitmasks == [1, 2, 3] and cog.pauli_indices == [0, 1]


def test_reference_count_mismatch_message():
    """test/test_cutting_reconstruction.py:139-150."""
    results = SamplerResult(quasi_dists=[QuasiDistribution({"0": 1.0})], metadata=[{}])
    with pytest.raises(ValueError) as e:
        oracle.reconstruct_literal(results, [(1.0, WeightType.EXACT)], PauliList(["ZZ", "XX"]))
    assert e.value.args[0] == KAT["count_mismatch_message"]


def test_reference_group_sizes():
    """test/utils/test_observable_grouping.py:104-107."""
    coll = ObservableCollection(PauliList(["XX", "ZZ", "XX"]))
    assert [len(g.commuting_observables) for g in coll.groups] == KAT["group_sizes_XX_ZZ_XX"] == [1, 1]


@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_literal_oracle_is_bit_identical_to_reference_source(name):
    case = _golden.load_case(name)
    out = oracle.reconstruct_literal(case["results"], case["coefficients"], case["observables"])
    assert [float(v).hex() for v in out] == [float(v).hex() for v in case["reference_expvals"]]


@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_exact_oracle_within_reference_self_noise(name):
    """E = tally/S differs from the reference's sequential accumulation only by its rounding noise."""
    case = _golden.load_case(name)
    out = oracle.reconstruct_exact(case["results"], case["coefficients"], case["observables"])
    kappa = sum(abs(c) for c, _ in case["coefficients"])
    assert np.max(np.abs(np.array(out) - np.array(case["reference_expvals"]))) <= 1e-13 * max(1.0, kappa)
    if name == "pow2":  # 1/S exact -> the reference has no rounding noise -> identical bits
        assert [float(v).hex() for v in out] == [float(v).hex() for v in case["reference_expvals"]]


def test_v1_case_bit_identical():
    case = _golden.load_v1_case()
    out = oracle.reconstruct_literal(case["results"], case["coefficients"], case["observables"])
    assert [float(v).hex() for v in out] == [float(v).hex() for v in case["reference_expvals"]]


@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_plan_builder_matches_reference_groups(name):
    """Vectorised masks / indices / lookup == what the reference's observable_grouping.py produced."""
    case = _golden.load_case(name)
    observables = case["observables"] if case["raw"]["partitioned"] else {"A": case["observables"]}
    for lab, (label, obs) in zip(case["raw"]["labels"], observables.items()):
        coll = ObservableCollection(obs)
        assert len(coll.groups) == len(lab["groups"])
        for cog, ref in zip(coll.groups, lab["groups"]):
            assert cog.general_observable.to_label() == ref["general"]
            assert [p.to_label() for p in cog.commuting_observables] == ref["members"]
            assert cog.pauli_indices == ref["pauli_indices"]
            assert [hex(m) for m in cog.pauli_bitmasks] == ref["pauli_bitmasks"]
        assert [[list(e) for e in slots] for slots in coll.lookup_indices] == lab["lookup"]
        # the oracle's own pure-Python grouping agrees too
        oc = oracle.build_collection(obs)
        assert [g["masks"] for g in oc["groups"]] == [cog.pauli_bitmasks for cog in coll.groups]


@pytest.mark.parametrize("name", ["tutorial", "wide_dense", "odd_widths", "molecular"])
def test_c_oracle_matches_python_oracle(name):
    """oracle_seq.c: sequential fp64 identical to the Python literal loop, tallies identical to the exact ones."""
    case = _golden.load_case(name)
    observables = case["observables"] if case["raw"]["partitioned"] else {"A": case["observables"]}
    results = case["results"] if case["raw"]["partitioned"] else {"A": case["results"]}
    for label, obs in observables.items():
        coll = ObservableCollection(obs)
        for idx, pub in enumerate(results[label]):
            cog = coll.groups[idx % len(coll.groups)]
            o, q = pub.data.observable_measurements.array, pub.data.qpd_measurements.array
            e_seq, tally = c_oracle.pub(o, q, cog.mask_bytes(o.shape[1]))
            assert np.array_equal(tally, oracle.tally_exact(o, q, cog.pauli_bitmasks))
            assert [v.hex() for v in e_seq] == [v.hex() for v in oracle.pub_expvals_literal(o, q, cog.pauli_bitmasks)]


def test_empty_and_edge_pubs():
    masks = [0, 1, 3]
    empty_o, empty_q = np.zeros((0, 1), np.uint8), np.zeros((0, 1), np.uint8)
    assert np.array_equal(oracle.tally_exact(empty_o, empty_q, masks), [0, 0, 0])
    assert np.array_equal(oracle.pub_expvals_literal(empty_o, empty_q, masks), [0.0, 0.0, 0.0])
    # qpd pad bits above num_bits still count (reference :213 uses every bit of the row)
    o = np.array([[1]], np.uint8)
    assert oracle.tally_exact(o, np.array([[0x80]], np.uint8), [1])[0] == 1
    assert oracle.tally_exact(o, np.array([[0x00]], np.uint8), [1])[0] == -1


def test_oracle_matches_the_reference_text_on_mixed_sampler_inputs():
    """SamplerV1 and SamplerV2 subsystems in one call (the reference decides per subsystem, :143,150): the literal oracle
    reproduces what the reference's own text returns, bit for bit."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference sources not available")
    import _cases

    results, coefficients, observables = _cases.case_mixed_samplers()
    ref = ref_loader.load_reference()["reconstruct_expectation_values"](results, coefficients, observables)
    got = oracle.reconstruct_literal(results, coefficients, observables)
    assert [float(v).hex() for v in got] == [float(v).hex() for v in ref]
