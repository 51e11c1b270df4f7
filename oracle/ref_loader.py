"""Loads the REFERENCE'S OWN source text as runnable code.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

Qiskit is not installable in this image, so ``import qiskit_addon_cutting`` fails at its first line.  The
hot path itself, however, only needs ``numpy`` plus a handful of Qiskit container classes, so this module

1. registers ``qiskit.quantum_info`` / ``qiskit.primitives`` shims whose ``Pauli`` / ``PauliList`` / result
   classes are the stand-ins of ``qiskit_addon_cutting_b200.containers``;
2. imports the reference's ``utils/observable_grouping.py`` (and ``utils/observable_terms.py``) unmodified from
   their files;
3. ``exec``s, unmodified, the text of ``reconstruct_expectation_values``, ``_process_outcome``,
   ``_process_outcome_v2``, ``_outcome_to_int`` (``cutting_reconstruction.py:31-232``), ``decompose_observables``
   (``cutting_decomposition.py:237-260``) and ``_get_pauli_indices`` (``cutting_experiments.py:362-370``).

Where the text comes from: ``oracle/_ref/qiskit_addon_cutting/`` when it exists (a git-ignored copy made by
``oracle/make_ref.py`` in the build container, so that it travels to the GPU box with the snapshot), else
``/root/reference/qiskit_addon_cutting`` (build container only).  Nothing of it is ever committed.

Users: ``tests/golden/make_golden.py`` (fixtures), ``bench.py --impl reference`` and ``bench.py``'s
``cpu_baseline`` leg (``kind: "reference"``), ``tests/test_oracle.py`` (oracle == reference on fresh inputs).
"""

from __future__ import annotations

import importlib.util
import sys
import types
from collections import defaultdict
from collections.abc import Hashable, Mapping, Sequence
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
VENDORED = HERE / "_ref" / "qiskit_addon_cutting"
UPSTREAM = Path("/root/reference/qiskit_addon_cutting")
# (relative path, first line, last line) of every piece that is executed; whole files are imported as modules
FILES = ["cutting_reconstruction.py", "cutting_decomposition.py", "cutting_experiments.py",
         "utils/observable_grouping.py", "utils/observable_terms.py"]


def reference_dir() -> Path | None:
    for cand in (VENDORED, UPSTREAM):
        if all((cand / f).exists() for f in FILES):
            return cand
    return None


def available() -> bool:
    return reference_dir() is not None


_loaded: dict | None = None


def load_reference() -> dict:
    """Namespace holding the reference's functions (``ns["reconstruct_expectation_values"]`` etc.)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    ref = reference_dir()
    if ref is None:
        raise FileNotFoundError("reference sources not found (run oracle/make_ref.py in the build container)")
    root = HERE.parent
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    from qiskit_addon_cutting_b200 import containers as C

    qi = types.ModuleType("qiskit.quantum_info")
    qi.Pauli, qi.PauliList, qi.SparsePauliOp = C.Pauli, C.PauliList, C.SparsePauliOp
    qp = types.ModuleType("qiskit.primitives")
    qp.SamplerResult, qp.PrimitiveResult = C.SamplerResult, C.PrimitiveResult
    qk = types.ModuleType("qiskit")
    qk.quantum_info, qk.primitives = qi, qp
    sys.modules.update({"qiskit": qk, "qiskit.quantum_info": qi, "qiskit.primitives": qp})

    spec = importlib.util.spec_from_file_location("_ref_observable_grouping", ref / "utils" / "observable_grouping.py")
    og = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = og  # dataclasses resolves annotations through sys.modules
    spec.loader.exec_module(og)

    def lines(path: Path, lo: int, hi: int) -> str:
        return "".join(path.read_text().splitlines(keepends=True)[lo - 1 : hi])

    ns = {
        "np": np,
        "Mapping": Mapping,
        "Sequence": Sequence,
        "Hashable": Hashable,
        "defaultdict": defaultdict,
        "PauliList": C.PauliList,
        "SamplerResult": C.SamplerResult,
        "PrimitiveResult": C.PrimitiveResult,
        "WeightType": C.WeightType,
        "CommutingObservableGroup": og.CommutingObservableGroup,
        "ObservableCollection": og.ObservableCollection,
        "observables_restricted_to_subsystem": og.observables_restricted_to_subsystem,
    }
    future = "from __future__ import annotations\n"
    exec(compile(future + lines(ref / "cutting_decomposition.py", 237, 260), "ref:decompose_observables", "exec"), ns)
    exec(compile(future + lines(ref / "cutting_experiments.py", 362, 370), "ref:_get_pauli_indices", "exec"), ns)
    exec(compile(future + lines(ref / "cutting_reconstruction.py", 31, 232), "ref:cutting_reconstruction", "exec"), ns)
    ns["_og"] = og
    ns["_source_dir"] = str(ref)
    _loaded = ns
    return ns


def load_reference_terms():
    """The reference's ``utils/observable_terms.py``, imported unmodified against the stand-ins."""
    load_reference()
    ref = reference_dir()
    spec = importlib.util.spec_from_file_location("_ref_observable_terms", ref / "utils" / "observable_terms.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod
