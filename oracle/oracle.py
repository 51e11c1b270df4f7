"""CPU oracle for the reconstruction hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``qiskit_addon_cutting_b200``) never does and fails loudly without its CUDA
library.

This file restates, in plain Python / NumPy, the algorithm of the reference
(paths relative to ``/root/reference/qiskit_addon_cutting``):

* :func:`outcome_to_int`            <- ``cutting_reconstruction.py:225-232``
* :func:`process_outcome_v2`        <- ``cutting_reconstruction.py:196-222``
* :func:`process_outcome`           <- ``cutting_reconstruction.py:173-193`` and
  ``cutting_experiments.py:362-370`` (identity workaround)
* :func:`group_masks`               <- ``utils/observable_grouping.py:57-113,134-157``
* :func:`build_collection`          <- ``utils/observable_grouping.py:168-227``
* :func:`reconstruct_literal`       <- ``cutting_reconstruction.py:79-170`` (validation, loop nest,
  sequential fp64 ``+= (1/shots) * sign`` accumulation, product and sum order)
* :func:`tally_exact`               <- the same sign rule, summed exactly as int64
* :func:`reconstruct_exact`         <- ``cutting_reconstruction.py:133-170`` with ``E = tally / shots``

Pinning: the oracle is checked against the reference's own known-answer tests
(``test/test_cutting_reconstruction.py:43-50,117-138,139-150,152-170``) and
against fixtures produced by executing the reference's *own source text*
(``tests/golden/make_golden.py``; fixtures committed under ``tests/golden/``).
What stays **parity unpinned**: the order of the commuting groups, which the
reference delegates to Qiskit / rustworkx (not vendored, not version-locked,
not installable here); see DESIGN.md.
"""

from __future__ import annotations

from collections import defaultdict
from collections.abc import Mapping

import numpy as np


# --------------------------------------------------------------------------------------
# Sign rule
# --------------------------------------------------------------------------------------
def outcome_to_int(outcome) -> int:
    """Parse a V1 outcome key: int, ``'0101'``, ``'0b0101'`` or ``'0x5'`` (spaces ignored)."""
    if isinstance(outcome, (int, np.integer)):
        return int(outcome)
    outcome = outcome.replace(" ", "")
    if len(outcome) < 2 or outcome[1] in ("0", "1"):
        return int(f"0b{outcome}", 0)
    return int(outcome, 0)


def process_outcome_v2(masks, obs_outcomes: int, qpd_outcomes: int) -> np.ndarray:
    """``(-1)**(popcount(qpd) + popcount(obs & mask_t))`` for every mask, as fp64 ±1."""
    qpd_factor = 1 - 2 * (int(qpd_outcomes).bit_count() & 1)
    rv = np.zeros(len(masks))
    for i, mask in enumerate(masks):
        obs = 1 - 2 * ((int(obs_outcomes) & mask).bit_count() & 1)
        rv[i] = qpd_factor * obs
    return rv


def process_outcome(masks, num_pauli_indices: int, outcome) -> np.ndarray:
    """V1 split: low ``max(1, num_pauli_indices)`` bits are the observable register."""
    num_meas_bits = max(1, num_pauli_indices)
    outcome = outcome_to_int(outcome)
    obs_outcomes = outcome & ((1 << num_meas_bits) - 1)
    qpd_outcomes = outcome >> num_meas_bits
    return process_outcome_v2(masks, obs_outcomes, qpd_outcomes)


# --------------------------------------------------------------------------------------
# Grouping (pure-Python loops over single-qubit Paulis, like the reference)
# --------------------------------------------------------------------------------------
def _zx(pauli) -> tuple[list[bool], list[bool]]:
    return [bool(v) for v in np.ravel(pauli.z)], [bool(v) for v in np.ravel(pauli.x)]


def group_masks(members) -> tuple[list[int], list[int]]:
    """Return ``(pauli_indices, pauli_bitmasks)`` of one commuting group."""
    if len(members) == 0:
        raise ValueError(
            "Empty input sequence: consider performing no experiments rather than an "
            "experiment over the identity Pauli."
        )
    n = len(_zx(members[0])[0])
    general = [(False, False)] * n
    for pauli in members:
        z, x = _zx(pauli)
        for i in range(n):
            o = (z[i], x[i])
            if o == (False, False) or general[i] == o:
                continue
            if general[i] != (False, False):
                raise ValueError(
                    "Observables are incompatible; cannot construct a single general observable."
                )
            general[i] = o
    pauli_indices = [i for i, g in enumerate(general) if g != (False, False)]
    masks = []
    for pauli in members:
        z, x = _zx(pauli)
        v = 0
        for i, j in enumerate(pauli_indices):
            if z[j] or x[j]:
                v |= 1 << i
        masks.append(v)
    return pauli_indices, masks


def _key(pauli) -> tuple:
    z, x = _zx(pauli)
    return (tuple(z), tuple(x))


def build_collection(observables) -> dict:
    """Groups, masks and lookup for one subsystem.

    Group membership/order comes from ``observables.unique().group_commuting(qubit_wise=True)``
    exactly as the reference asks Qiskit for it (``utils/observable_grouping.py:176,186``).
    """
    unique = observables.unique()
    groups = []
    lookup = defaultdict(list)
    for gi, group in enumerate(unique.group_commuting(qubit_wise=True)):
        members = list(group)
        pauli_indices, masks = group_masks(members)
        groups.append({"pauli_indices": pauli_indices, "masks": masks, "members": members})
        for ti, member in enumerate(members):
            lookup[_key(member)].append((gi, ti))
    return {"groups": groups, "lookup": dict(lookup)}


# --------------------------------------------------------------------------------------
# Reconstruction
# --------------------------------------------------------------------------------------
def _is_pauli_list(obj) -> bool:
    return (
        not isinstance(obj, Mapping)
        and hasattr(obj, "z")
        and hasattr(obj, "x")
        and hasattr(obj, "phase")
        and np.ndim(getattr(obj, "z")) == 2
    )


def _is_result(obj) -> bool:
    return hasattr(obj, "quasi_dists") or (
        not isinstance(obj, (Mapping, list, tuple)) and hasattr(obj, "__len__") and hasattr(obj, "__getitem__")
    )


def _normalise(results, observables):
    if _is_pauli_list(observables):
        if not _is_result(results):
            raise ValueError(
                "If observables is a PauliList, results must be a SamplerResult or PrimitiveResult instance."
            )
        if any(obs.phase != 0 for obs in observables):
            raise ValueError("An input observable has a phase not equal to 1.")
        return {"A": observables}, {"A": results}, len(observables)
    if isinstance(observables, Mapping):
        if not isinstance(results, Mapping):
            raise ValueError("If observables is a dictionary, results must also be a dictionary.")
        if observables.keys() != results.keys():
            raise ValueError("The subsystem labels of the observables and results do not match.")
        for subobservable in observables.values():
            if any(obs.phase != 0 for obs in subobservable):
                raise ValueError("An input observable has a phase not equal to 1.")
        return observables, results, len(list(observables.values())[0])
    raise ValueError("observables must be either a PauliList or dict.")


def _validated_collections(results_dict, coefficients, subobservables):
    collections = {label: build_collection(obs) for label, obs in subobservables.items()}
    for label, coll in collections.items():
        current = results_dict[label]
        if hasattr(current, "quasi_dists"):
            current = current.quasi_dists
        if len(current) != len(coefficients) * len(coll["groups"]):
            raise ValueError(
                f"The number of subexperiments performed in subsystem '{label}' "
                f"({len(current)}) should equal the number of coefficients "
                f"({len(coefficients)}) times the number of mutually commuting "
                f"subobservable groups ({len(coll['groups'])}), but it does not."
            )
    return collections


def pub_expvals_literal(obs_array, qpd_array, masks) -> np.ndarray:
    """One PUB, literal semantics: per-shot ``int.from_bytes`` and sequential fp64 ``+=``."""
    shots = qpd_array.shape[0]
    acc = np.zeros(len(masks))
    for j in range(shots):
        obs_outcomes = int.from_bytes(obs_array[j], "big")
        qpd_outcomes = int.from_bytes(qpd_array[j], "big")
        acc += (1 / shots) * process_outcome_v2(masks, obs_outcomes, qpd_outcomes)
    return acc


def tally_exact(obs_array, qpd_array, masks) -> np.ndarray:
    """One PUB, exact: ``tally[t] = sum_j sign_j(t)`` as int64 (vectorised over shots)."""
    shots = qpd_array.shape[0]
    obs_array = np.ascontiguousarray(obs_array[:shots])
    nb = obs_array.shape[1]
    qpd_par = np.bitwise_count(np.ascontiguousarray(qpd_array)).sum(axis=1, dtype=np.int64) & 1
    out = np.empty(len(masks), dtype=np.int64)
    for t, mask in enumerate(masks):
        mbytes = np.frombuffer(int(mask & ((1 << (8 * nb)) - 1)).to_bytes(nb, "big"), dtype=np.uint8)
        par = (np.bitwise_count(obs_array & mbytes[None, :]).sum(axis=1, dtype=np.int64) + qpd_par) & 1
        out[t] = shots - 2 * int(par.sum())
    return out


def _reconstruct(results, coefficients, observables, pub_fn, v1_fn):
    subobservables, results_dict, num_obs = _normalise(results, observables)
    expvals = np.zeros(num_obs)
    collections = _validated_collections(results_dict, coefficients, subobservables)
    for i, coeff in enumerate(coefficients):
        current_expvals = np.ones((num_obs,))
        for label, coll in collections.items():
            groups = coll["groups"]
            subsystem_expvals = []
            current_result = results_dict[label]
            for k, group in enumerate(groups):
                idx = i * len(groups) + k
                if hasattr(current_result, "quasi_dists"):
                    subsystem_expvals.append(v1_fn(current_result.quasi_dists[idx], group))
                else:
                    data_pub = current_result[idx].data
                    obs_array = data_pub.observable_measurements.array
                    qpd_array = data_pub.qpd_measurements.array
                    subsystem_expvals.append(pub_fn(obs_array, qpd_array, group["masks"]))
            for k, subobservable in enumerate(subobservables[label]):
                current_expvals[k] *= np.mean(
                    [subsystem_expvals[m][n] for m, n in coll["lookup"][_key(subobservable)]]
                )
        expvals += coeff[0] * current_expvals
    return list(expvals)


def _v1_literal(quasi_probs, group) -> np.ndarray:
    acc = np.zeros(len(group["masks"]))
    for outcome, quasi_prob in quasi_probs.items():
        acc += quasi_prob * process_outcome(group["masks"], len(group["pauli_indices"]), outcome)
    return acc


def reconstruct_literal(results, coefficients, observables) -> list:
    """Reference semantics, including the sequential fp64 rounding of every shot's ``±(1/S)``."""
    return _reconstruct(results, coefficients, observables, pub_expvals_literal, _v1_literal)


def reconstruct_exact(results, coefficients, observables) -> list:
    """Same loop nest with per-PUB ``E = tally / shots`` (one correctly rounded division)."""

    def pub_fn(obs_array, qpd_array, masks):
        shots = qpd_array.shape[0]
        if shots == 0:  # the reference's shot loop does not execute: expectation values stay 0
            return np.zeros(len(masks))
        return tally_exact(obs_array, qpd_array, masks) / shots

    return _reconstruct(results, coefficients, observables, pub_fn, _v1_literal)


def all_tallies(results, coefficients, observables) -> dict:
    """``{label: [tally array per PUB]}`` for V2 results — the bit-exact target of kernel K1."""
    subobservables, results_dict, _ = _normalise(results, observables)
    collections = _validated_collections(results_dict, coefficients, subobservables)
    out = {}
    for label, coll in collections.items():
        groups = coll["groups"]
        result = results_dict[label]
        tallies = []
        for idx in range(len(coefficients) * len(groups)):
            data_pub = result[idx].data
            tallies.append(
                tally_exact(
                    data_pub.observable_measurements.array,
                    data_pub.qpd_measurements.array,
                    groups[idx % len(groups)]["masks"],
                )
            )
        out[label] = tallies
    return out


# ---- term recombination (reference utils/observable_terms.py:22-90) ------------------------------------
_PHASE_FACTOR = (1 + 0j, -1j, -1 + 0j, 1j)  # (-i)**phase


def _op_terms(observable):
    """(phase-free key, coefficient) pairs of SparsePauliOp(observable) (reference :33-35,65-67)."""
    if hasattr(observable, "paulis"):
        return [(_key(p), c) for p, c in zip(observable.paulis, observable.coeffs, strict=True)]
    return [(_key(observable), np.complex128(_PHASE_FACTOR[int(observable.phase) % 4]))]


def unique_term_keys(observables) -> list:
    """Keys of ``gather_unique_observable_terms`` in its order (reference :30-41): first appearance, zero
    coefficients skipped."""
    seen, out = set(), []
    for observable in observables:
        for key, coeff in _op_terms(observable):
            if coeff == 0:
                continue
            if key not in seen:
                seen.add(key)
                out.append(key)
    return out


def observable_expvals_from_terms(observables, term_expvals: Mapping) -> list:
    """Literal restatement of reference :60-90: ``rv = 0.0j; rv += coeff * term_expval`` per term, in order.

    ``term_expvals`` maps Paulis (phase-free) to numbers; a missing term raises the reference's ValueError.
    """
    table = {_key(p): v for p, v in term_expvals.items()}
    out = []
    for observable in observables:
        rv = 0.0j
        for key, coeff in _op_terms(observable):
            if coeff == 0:
                continue
            if key not in table:
                raise ValueError("An observable contains a Pauli term whose expectation value was not provided.")
            rv += coeff * table[key]
        out.append(rv)
    return out
