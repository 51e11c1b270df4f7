"""Validation order and exact error strings of the drop-in API (reference test_cutting_reconstruction.py:51-116,139-150).

All of these raise before any device work, so they run without a GPU.
"""

import numpy as np
import pytest

from qiskit_addon_cutting_b200 import (
    PauliList, QuasiDistribution, SamplerResult, WeightType, reconstruct_expectation_values,
)


def _v1():
    return SamplerResult(quasi_dists=[QuasiDistribution({"0": 1.0})], metadata=[{}])


def test_mismatching_input_types():
    weights = [(0.5, WeightType.EXACT), (0.5, WeightType.EXACT)]
    observables = {"A": PauliList(["Z"]), "B": PauliList(["Z"])}
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values(_v1(), weights, observables)
    assert e.value.args[0] == "If observables is a dictionary, results must also be a dictionary."
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values({"A": _v1()}, weights, PauliList(["ZZ"]))
    assert (
        e.value.args[0]
        == "If observables is a PauliList, results must be a SamplerResult or PrimitiveResult instance."
    )


def test_invalid_observables_type():
    class SparsePauliOpLike:  # a list of operator objects is neither a PauliList nor a dict
        pass

    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values(_v1(), [(1.0, WeightType.EXACT)], [SparsePauliOpLike()])
    assert e.value.args[0] == "observables must be either a PauliList or dict."


def test_mismatching_subsystem_labels():
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values({"A": _v1()}, [(1.0, WeightType.EXACT)], {"B": [PauliList("ZZ")]})
    assert e.value.args[0] == "The subsystem labels of the observables and results do not match."


def test_unsupported_phase():
    weights = [(0.5, WeightType.EXACT)]
    observables = PauliList(["iZZ"])
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values(_v1(), weights, observables)
    assert e.value.args[0] == "An input observable has a phase not equal to 1."
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values({"A": _v1()}, weights, {"A": observables})
    assert e.value.args[0] == "An input observable has a phase not equal to 1."


def test_inconsistent_number_of_subexperiments():
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values(_v1(), [(1.0, WeightType.EXACT)], PauliList(["ZZ", "XX"]))
    assert (
        e.value.args[0]
        == "The number of subexperiments performed in subsystem 'A' (1) should equal the number of coefficients (1) times the number of mutually commuting subobservable groups (2), but it does not."
    )


def test_type_check_precedes_key_check_precedes_phase_check():
    # dict observables with a phase AND mismatching keys -> the key error wins (reference :95-106)
    with pytest.raises(ValueError) as e:
        reconstruct_expectation_values({"A": _v1()}, [(1.0, WeightType.EXACT)], {"B": PauliList(["iZ"])})
    assert e.value.args[0] == "The subsystem labels of the observables and results do not match."


def test_no_cpu_fallback(monkeypatch):
    """Valid inputs without a CUDA device must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qiskit_addon_cutting_b200._lib import QcutError

    with pytest.raises((QcutError, ImportError)):
        reconstruct_expectation_values(_v1(), [(1.0, WeightType.EXACT)], PauliList(["ZZ"]))


def test_product_never_imports_the_oracle():
    import pathlib
    import re

    pkg = pathlib.Path(__file__).resolve().parents[1] / "qiskit_addon_cutting_b200"
    for path in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if path.suffix in (".py", ".cu", ".cuh", ".h"):
            assert not re.search(r"^\s*(from|import)\s+oracle|oracle_seq", path.read_text(), re.M), path


def test_no_coefficients_returns_zeros_like_the_reference():
    """coefficients=[] with empty results: the reference's loops do not execute and it returns zeros
    (cutting_reconstruction.py:112,133-170); no device is needed for that."""
    from oracle import oracle
    from qiskit_addon_cutting_b200 import PauliList, PrimitiveResult, reconstruct_expectation_values

    observables = {"A": PauliList(["ZI", "IZ", "XX"]), "B": PauliList(["ZZ", "XI", "IX"])}
    results = {"A": PrimitiveResult([]), "B": PrimitiveResult([])}
    got = reconstruct_expectation_values(results, [], observables)
    assert got == [0.0, 0.0, 0.0] and all(isinstance(v, np.float64) for v in got)
    assert got == oracle.reconstruct_literal(results, [], observables)
    assert reconstruct_expectation_values(PrimitiveResult([]), [], PauliList(["ZZ", "XX"])) == [0.0, 0.0]
