"""B200-native reconstruction for qiskit-addon-cutting (drop-in ``reconstruct_expectation_values``)."""

from .containers import (
    BitArray,
    DataBin,
    Pauli,
    PauliList,
    PrimitiveResult,
    PubResult,
    QuasiDistribution,
    SamplerResult,
    SparsePauliOp,
    WeightType,
)
from .cutting_reconstruction import reconstruct_expectation_values
from .observable_grouping import CommutingObservableGroup, ObservableCollection
from .observable_terms import (
    TermRecombination,
    gather_unique_observable_terms,
    reconstruct_observable_expvals_from_terms,
)

__all__ = [
    "reconstruct_expectation_values",
    "install",
    "gather_unique_observable_terms",
    "reconstruct_observable_expvals_from_terms",
    "TermRecombination",
    "SparsePauliOp",
    "CommutingObservableGroup",
    "ObservableCollection",
    "WeightType",
    "Pauli",
    "PauliList",
    "BitArray",
    "DataBin",
    "PubResult",
    "PrimitiveResult",
    "QuasiDistribution",
    "SamplerResult",
]

__version__ = "0.1.0"


def install() -> bool:
    """Patch ``qiskit_addon_cutting.reconstruct_expectation_values`` with the B200 path.

    Returns ``False`` (and does nothing) when the reference package is not importable.
    """
    try:
        import qiskit_addon_cutting
        from qiskit_addon_cutting import cutting_reconstruction as ref_module
    except ImportError:
        return False
    ref_module.reconstruct_expectation_values = reconstruct_expectation_values
    qiskit_addon_cutting.reconstruct_expectation_values = reconstruct_expectation_values
    return True
