"""The drop-in against REAL Qiskit objects.  Qiskit is not installable in the build image (no wheels, no network), so
these tests skip here; they are for whoever has ``qiskit`` (and, for the workflow test, ``qiskit_addon_cutting`` and
a local sampler) next to a B200.  They mirror what the reference tests do with the reference function:
``test/test_cutting_reconstruction.py:117-138`` (SamplerV2 known answer), ``:139-150`` (count check) and
``test/test_cutting_workflows.py:186-219`` (generate -> sample -> reconstruct), plus ``install()``.

What they pin that the stand-in based tests cannot: ``ObservableCollection`` calls the CALLER's
``PauliList.unique().group_commuting(qubit_wise=True)`` (reference ``utils/observable_grouping.py:176,186``), i.e.
Qiskit / rustworkx decide the order of the commuting groups, and the result walk reads real ``PrimitiveResult`` /
``DataBin`` / ``BitArray`` attributes."""

import numpy as np
import pytest

qiskit = pytest.importorskip("qiskit")
pytestmark = pytest.mark.gpu


def _v2_golden_inputs():
    from qiskit.primitives import BitArray, PrimitiveResult, PubResult
    from qiskit.primitives.containers import make_data_bin
    from qiskit.quantum_info import PauliList

    data_bin_cls = make_data_bin([("observable_measurements", BitArray), ("qpd_measurements", BitArray)], shape=())
    obs = BitArray(np.array([0, 0, 0, 1, 0, 0, 2, 3, 1, 0], dtype=np.uint8).reshape(-1, 1), 2)
    qpd = BitArray(np.array([0, 1, 1, 3, 2, 0, 1, 3, 0, 2], dtype=np.uint8).reshape(-1, 1), 2)
    results = PrimitiveResult([PubResult(data_bin_cls(observable_measurements=obs, qpd_measurements=qpd))])
    return results, PauliList(["II", "IZ", "ZI", "ZZ"])


def test_samplerv2_known_answer_with_qiskit_containers(lib):
    from qiskit_addon_cutting_b200 import WeightType, reconstruct_expectation_values

    results, observables = _v2_golden_inputs()
    out = reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], observables)
    assert out == pytest.approx([0.0, -0.6, 0.0, -0.2])
    assert reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], observables, rounding="reference") == \
        pytest.approx([0.0, -0.6, 0.0, -0.2], abs=1e-15)


def test_count_check_and_phase_errors_with_qiskit_containers(lib):
    from qiskit.primitives import SamplerResult
    from qiskit.quantum_info import PauliList
    from qiskit.result import QuasiDistribution

    from qiskit_addon_cutting_b200 import WeightType, reconstruct_expectation_values

    results = SamplerResult(quasi_dists=[QuasiDistribution({"0": 1.0})], metadata=[{}])
    with pytest.raises(ValueError) as e_info:
        reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], PauliList(["ZZ", "XX"]))
    assert e_info.value.args[0] == (
        "The number of subexperiments performed in subsystem 'A' (1) should equal the number of coefficients (1) "
        "times the number of mutually commuting subobservable groups (2), but it does not."
    )
    with pytest.raises(ValueError) as e_info:
        reconstruct_expectation_values(results, [(0.5, WeightType.EXACT)], PauliList(["iZZ"]))
    assert e_info.value.args[0] == "An input observable has a phase not equal to 1."
    assert reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], PauliList(["ZZ"])) == [1.0]


def test_grouping_is_qiskits_own(lib):
    """The plan's groups are the ones Qiskit's ``group_commuting`` returns, in its order (what
    ``generate_cutting_experiments`` used when it emitted the subexperiments)."""
    from qiskit.quantum_info import PauliList

    from qiskit_addon_cutting_b200 import ObservableCollection

    paulis = PauliList(["IIZ", "IZI", "XII", "IXX", "ZZZ", "YIY", "IIZ"])
    ours = ObservableCollection(paulis)
    want = paulis.unique().group_commuting(qubit_wise=True)
    assert [sorted(g.commuting_observables.to_labels()) for g in ours.groups] == [sorted(g.to_labels()) for g in want]


def test_install_patches_the_reference_and_matches_it(lib):
    """With the reference importable: same inputs through the reference's own function and through ours."""
    ref_pkg = pytest.importorskip("qiskit_addon_cutting")
    from qiskit_addon_cutting.qpd import WeightType as RefWeightType

    import qiskit_addon_cutting_b200 as ours

    results, observables = _v2_golden_inputs()
    want = ref_pkg.reconstruct_expectation_values(results, [(1.0, RefWeightType.EXACT)], observables)
    original = ref_pkg.reconstruct_expectation_values
    try:
        assert ours.install() is True
        assert ref_pkg.reconstruct_expectation_values is ours.reconstruct_expectation_values
        got = ref_pkg.reconstruct_expectation_values(results, [(1.0, RefWeightType.EXACT)], observables)
        lit = ref_pkg.reconstruct_expectation_values(results, [(1.0, RefWeightType.EXACT)], observables, rounding="reference")
    finally:
        ref_pkg.reconstruct_expectation_values = original
        ref_pkg.cutting_reconstruction.reconstruct_expectation_values = original
    assert got == pytest.approx(want, abs=1e-12)
    assert [float(v).hex() for v in lit] == [float(v).hex() for v in want]


def test_workflow_generate_sample_reconstruct(lib):
    """The tutorial workflow end to end (reference test_cutting_workflows.py:186-219), reference result vs ours on the
    same sampler output."""
    ref_pkg = pytest.importorskip("qiskit_addon_cutting")
    from qiskit import QuantumCircuit
    from qiskit.primitives import StatevectorSampler
    from qiskit.quantum_info import PauliList

    import qiskit_addon_cutting_b200 as ours

    qc = QuantumCircuit(4)
    qc.h(range(4))
    qc.ryy(0.1, 1, 2)
    qc.h(range(4))
    observables = PauliList(["ZIII", "IZII", "IIZI", "IIIZ"])
    problem = ref_pkg.partition_problem(circuit=qc, partition_labels=[0, 0, 1, 1], observables=observables)
    subexperiments, coefficients = ref_pkg.generate_cutting_experiments(
        circuits=problem.subcircuits, observables=problem.subobservables, num_samples=100)
    sampler = StatevectorSampler(seed=7)
    results = {label: sampler.run(subexperiments[label], shots=128).result() for label in subexperiments}
    want = ref_pkg.reconstruct_expectation_values(results, coefficients, problem.subobservables)
    got = ours.reconstruct_expectation_values(results, coefficients, problem.subobservables)
    lit = ours.reconstruct_expectation_values(results, coefficients, problem.subobservables, rounding="reference")
    assert len(got) == len(observables)
    assert np.max(np.abs(np.array(got) - np.array(want))) <= 1e-12
    assert [float(v).hex() for v in lit] == [float(v).hex() for v in want]
