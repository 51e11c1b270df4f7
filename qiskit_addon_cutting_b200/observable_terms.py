"""Unique Pauli terms of a set of observables, and their recombination on the GPU.

Mirror of the reference's ``qiskit_addon_cutting/utils/observable_terms.py`` -- the step on either side of
the hot path (``SURVEY.md`` section 8f, row 3):

* :func:`gather_unique_observable_terms` (reference ``:22-57``) is pure bookkeeping on Python objects and
  stays on the host, with the reference's semantics and error strings;
* :func:`reconstruct_observable_expvals_from_terms` (reference ``:60-90``) is the arithmetic: every
  observable's value ``sum_j coeff_j * <P_j>``.  The dict lookups (and the ``ValueError`` for a missing
  term) happen on the host; the complex fp64 sums run in ``qcut_recombine_launch`` (``csrc/reduce.cu``),
  sequentially per observable, so results equal the reference's bit for bit.
  :class:`TermRecombination` is the same computation on term values that are still in HBM (the output
  of ``reconstruct_expectation_values``), without a round trip through the host.

Real Qiskit objects are accepted by protocol: an operator has ``.paulis`` / ``.coeffs``, a Pauli has
one-dimensional ``.z`` / ``.x`` and ``.phase``.
"""

from __future__ import annotations

from collections.abc import Iterable, Mapping, Sequence

import numpy as np

from . import containers

_PHASE_FACTOR = (1 + 0j, -1j, -1 + 0j, 1j)  # (-i)**phase


def _is_pauli(obj) -> bool:
    return all(hasattr(obj, a) for a in ("z", "x", "phase")) and np.ndim(obj.z) == 1


def _is_pauli_list(obj) -> bool:
    return all(hasattr(obj, a) for a in ("z", "x", "phase")) and np.ndim(obj.z) == 2


def _terms_of(observable):
    """``zip(op.paulis, op.coeffs)`` of ``SparsePauliOp(observable)`` (reference ``:33-35,65-67``)."""
    if _is_pauli(observable):
        bare = type(observable)((np.asarray(observable.z), np.asarray(observable.x)))  # the phase moves to the coefficient
        return [(bare, np.complex128(_PHASE_FACTOR[int(observable.phase) % 4]))]
    if hasattr(observable, "paulis") and hasattr(observable, "coeffs"):
        paulis, coeffs = observable.paulis, observable.coeffs
        if len(paulis) != len(coeffs):
            raise ValueError("zip() argument 2 is shorter than argument 1")  # zip(..., strict=True) in the reference
        return list(zip(paulis, coeffs))
    raise TypeError(f"Unsupported observable type: {type(observable).__name__}")


def gather_unique_observable_terms(observables):
    """Inspect the contents of each observable to find and return the unique Pauli terms.

    Reference ``utils/observable_terms.py:22-57``: first-appearance order, zero coefficients skipped,
    phases dropped.  Returns a ``PauliList`` of the caller's flavour (Qiskit's when given Qiskit objects).
    """
    if len(observables) == 0:
        if _is_pauli_list(observables):
            return observables.copy()
        raise ValueError("observables list cannot be empty")

    pauli_list: list = []
    pauli_set: set = set()
    for observable in observables:
        for pauli, coeff in _terms_of(observable):
            if coeff == 0:
                continue
            if pauli not in pauli_set:
                pauli_list.append(pauli)
                pauli_set.add(pauli)

    list_type = _pauli_list_type(observables, pauli_list)
    if not pauli_list:
        # every coefficient is zero: an empty list with the right number of qubits (reference :46-51)
        return list_type(["I" * observables[0].num_qubits])[:0]
    try:
        return list_type(pauli_list)
    except ValueError as ex:
        raise ValueError(
            "Cannot construct PauliList.  Do provided observables all have "
            "the same number of qubits?"
        ) from ex


def _pauli_list_type(observables, paulis):
    """The ``PauliList`` class that goes with the caller's objects (ours, or Qiskit's when it is in use)."""
    if _is_pauli_list(observables):
        return type(observables)
    probe = paulis[0] if paulis else None
    if probe is None:
        first = observables[0]
        if hasattr(first, "paulis"):
            return type(first.paulis)
        probe = first
    if isinstance(probe, containers.Pauli):
        return containers.PauliList
    from importlib import import_module

    return import_module(type(probe).__module__.split(".")[0] + ".quantum_info").PauliList  # qiskit.quantum_info


class TermRecombination:
    """Rows of ``(term index, coefficient)`` per observable, resident in HBM (CSR), for ``qcut_recombine_launch``.

    ``term_index`` maps a phase-free Pauli to its position in the vector of term expectation values.
    """

    def __init__(self, observables: Iterable, term_index: Mapping):
        row_start, index, coeffs = [0], [], []
        for observable in observables:
            for pauli, coeff in _terms_of(observable):
                if coeff == 0:
                    continue
                try:
                    index.append(term_index[pauli])
                except KeyError as ex:
                    raise ValueError(
                        "An observable contains a Pauli term whose expectation value "
                        "was not provided."
                    ) from ex
                coeffs.append(complex(coeff))
            row_start.append(len(index))
        self.n_rows = len(row_start) - 1
        self.row_start = np.array(row_start, dtype=np.uint32)
        self.index = np.array(index, dtype=np.uint32)
        self.coeffs = np.array(coeffs, dtype=np.complex128)
        self._device = None

    def _tables(self):
        if self._device is None:
            from . import engine

            self._device = (
                engine._to_device(self.row_start),
                engine._to_device(self.index if len(self.index) else np.zeros(1, dtype=np.uint32)),
                engine._to_device(self.coeffs.view(np.float64) if len(self.coeffs) else np.zeros(2)),
            )
        return self._device

    def launch(self, term_expvals):
        """``term_expvals``: CUDA tensor, float64 or complex128.  Returns a complex128 CUDA tensor [n_rows]."""
        import torch

        from . import _lib, engine

        _lib.require_device()
        if term_expvals.dtype not in (torch.float64, torch.complex128) or not term_expvals.is_cuda:
            raise ValueError("term_expvals must be a float64 or complex128 CUDA tensor")
        term_expvals = term_expvals.contiguous()
        out = torch.empty(max(self.n_rows, 1), dtype=torch.complex128, device="cuda")
        row_start, index, coeffs = self._tables()
        _lib.check(
            _lib.load().qcut_recombine_launch(
                term_expvals.data_ptr(), int(term_expvals.dtype == torch.complex128), row_start.data_ptr(),
                index.data_ptr(), coeffs.data_ptr(), self.n_rows, out.data_ptr(), engine._stream_ptr(),
            ),
            "qcut_recombine_launch",
        )
        return out[: self.n_rows]


def reconstruct_observable_expvals_from_terms(observables, term_expvals: Mapping) -> list[complex]:
    """Reconstruct the expectation values given the expectation value of each unique term.

    Reference ``utils/observable_terms.py:82-90``.  Raises ``ValueError`` ("An observable contains a Pauli
    term whose expectation value was not provided.") like the reference; needs the CUDA library and a device
    as soon as there is at least one observable -- there is no CPU fallback.
    """
    keys = list(term_expvals)
    recombination = TermRecombination(observables, {pauli: i for i, pauli in enumerate(keys)})
    if recombination.n_rows == 0:
        return []
    from . import _lib

    _lib.require_device()
    import torch

    values = np.array([term_expvals[k] for k in keys] if keys else [0.0])
    values = values.astype(np.complex128 if np.iscomplexobj(values) else np.float64)
    out = recombination.launch(torch.from_numpy(values).cuda())
    return list(out.cpu().numpy())
