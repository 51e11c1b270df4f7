"""Plan builder: commuting observable groups and their bitmasks, vectorised on symplectic arrays.

Host-side mirror of the part of ``qiskit_addon_cutting/utils/observable_grouping.py``
that the reconstruction hot path consumes (``SURVEY.md`` §8 a3):

* ``CommutingObservableGroup.pauli_indices`` / ``pauli_bitmasks`` — reference
  ``utils/observable_grouping.py:134-157``;
* ``ObservableCollection.groups`` / ``lookup`` — reference ``:168-227``;
* ``most_general_observable`` — reference ``:57-113``.

The reference walks Python ``Pauli`` objects qubit by qubit (its own
``TODO(perf)`` notes at ``:89-93,136``); here every step is a NumPy operation on
the ``(terms, qubits)`` boolean ``z`` / ``x`` matrices, and the masks are
additionally exported in the packed big-endian byte layout the CUDA kernels AND
against raw ``BitArray`` rows.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .containers import Pauli, PauliList


def _symplectic(observables) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Return ``(z, x, phase)`` of a PauliList-like object or an iterable of Pauli-likes."""
    z = getattr(observables, "z", None)
    x = getattr(observables, "x", None)
    if z is not None and x is not None and np.ndim(z) == 2:
        phase = np.asarray(getattr(observables, "phase", np.zeros(len(z), dtype=np.int64)))
        return np.asarray(z, dtype=bool), np.asarray(x, dtype=bool), np.atleast_1d(phase)
    paulis = list(observables)
    if not paulis:
        raise ValueError("Empty observable list.")
    zs = np.array([np.asarray(p.z, dtype=bool).ravel() for p in paulis])
    xs = np.array([np.asarray(p.x, dtype=bool).ravel() for p in paulis])
    ph = np.array([int(np.ravel(getattr(p, "phase", 0))[0]) for p in paulis], dtype=np.int64)
    return zs, xs, ph


def most_general_observable(commuting_observables, /, num_qubits: int | None = None) -> Pauli:
    """Smallest Pauli string whose measurement determines every member (reference ``:57-113``)."""
    if len(commuting_observables) == 0:
        raise ValueError(
            "Empty input sequence: consider performing no experiments rather than an "
            "experiment over the identity Pauli."
        )
    for j, obs in enumerate(commuting_observables):
        if not (hasattr(obs, "z") and hasattr(obs, "x")):
            raise ValueError("Input sequence includes something other than a Pauli.")
        if num_qubits is None:
            num_qubits = len(obs)
        if len(obs) != num_qubits:
            raise ValueError(
                f"Observable {j} has incorrect qubit count ({len(obs)} rather than {num_qubits})."
            )
    z, x, _ = _symplectic(commuting_observables)
    gz, gx = _general_observable_zx(z, x)
    return Pauli((gz, gx))


def _general_observable_zx(z: np.ndarray, x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    non_identity = z | x
    gz = (z & non_identity).any(axis=0)
    gx = (x & non_identity).any(axis=0)
    # Every non-identity entry must agree with the general observable on its qubit.
    clash = non_identity & ((z != gz[None, :]) | (x != gx[None, :]))
    if clash.any():
        raise ValueError(
            "Observables are incompatible; cannot construct a single general observable."
        )
    return gz, gx


def masks_from_symplectic(z: np.ndarray, x: np.ndarray, pauli_indices: np.ndarray) -> list[int]:
    """Bit ``i`` of mask ``t`` is set iff member ``t`` is non-identity on ``pauli_indices[i]``."""
    if len(pauli_indices) == 0:
        return [0] * z.shape[0]
    bits = (z | x)[:, pauli_indices]
    packed = np.packbits(bits, axis=1, bitorder="little")
    return [int.from_bytes(row.tobytes(), "little") for row in packed]


@dataclass(frozen=True)
class CommutingObservableGroup:
    """Set of mutually qubit-wise commuting observables (reference ``:116-157``)."""

    general_observable: Pauli
    commuting_observables: list
    pauli_indices: list[int] = field(init=False)
    pauli_bitmasks: list[int] = field(init=False)

    def __post_init__(self) -> None:
        z, x, phase = _symplectic(self.commuting_observables)
        bad = np.flatnonzero(phase != 0)
        if bad.size:
            raise ValueError(
                "CommutingObservableGroup only supports Paulis with phase == 0. "
                f"(Value provided: {int(phase[bad[0]])})"
            )
        gz = np.asarray(self.general_observable.z, dtype=bool).ravel()
        gx = np.asarray(self.general_observable.x, dtype=bool).ravel()
        pauli_indices = np.flatnonzero(gz | gx)
        object.__setattr__(self, "pauli_indices", [int(i) for i in pauli_indices])
        object.__setattr__(self, "pauli_bitmasks", masks_from_symplectic(z, x, pauli_indices))

    @property
    def num_measured_bits(self) -> int:
        """Width of ``observable_measurements``: ``len(_get_pauli_indices(cog))``.

        Identity-only groups still measure qubit 0 into a 1-bit register
        (reference ``cutting_experiments.py:362-370``).
        """
        return max(1, len(self.pauli_indices))

    def mask_bytes(self, nbytes: int) -> np.ndarray:
        """Masks as ``(T, nbytes)`` uint8, big-endian rows, truncated/padded to ``nbytes``.

        This is the memory order of ``BitArray.array`` rows, so the kernels AND
        raw bytes without byte swapping (popcount is byte-order agnostic).
        """
        width = (1 << (8 * nbytes)) - 1
        rows = [int(m & width).to_bytes(nbytes, "big") for m in self.pauli_bitmasks]
        return np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), nbytes).copy()


class ObservableCollection:
    """Observables organised into qubit-wise commuting groups (reference ``:160-261``).

    ``lookup_indices[k]`` lists the ``(group, term)`` slots of input observable
    ``k`` — the array form of the reference's ``lookup[Pauli]`` dict
    (``:220-224``), ready to be shipped to the device.
    """

    def __init__(self, observables, /):
        z, x, phase = _symplectic(observables)
        self._num_qubits = z.shape[1]

        # unique(): first occurrences in original order (reference :175-178).
        if hasattr(observables, "unique") and hasattr(observables, "group_commuting"):
            unique = observables.unique()
            groups_zx = [
                (np.asarray(g.z, dtype=bool), np.asarray(g.x, dtype=bool), g)
                for g in unique.group_commuting(qubit_wise=True)
            ]
        else:
            unique = PauliList.from_symplectic(z, x, phase).unique()
            groups_zx = [(g.z, g.x, g) for g in unique.group_commuting(qubit_wise=True)]

        groups: list[CommutingObservableGroup] = []
        slot: dict[bytes, list[tuple[int, int]]] = {}
        for gi, (gz_rows, gx_rows, members) in enumerate(groups_zx):
            gz, gx = _general_observable_zx(gz_rows, gx_rows)
            general = self.construct_general_observable(gz, gx)
            groups.append(CommutingObservableGroup(general, list(members)))
            for ti in range(gz_rows.shape[0]):
                key = np.concatenate([gz_rows[ti], gx_rows[ti]]).tobytes()
                slot.setdefault(key, []).append((gi, ti))
        self._groups = groups
        self._slot = slot
        self.lookup_indices: list[list[tuple[int, int]]] = [
            slot[np.concatenate([z[k], x[k]]).tobytes()] for k in range(z.shape[0])
        ]

    @staticmethod
    def construct_general_observable(gz: np.ndarray, gx: np.ndarray) -> Pauli:
        return Pauli((gz, gx))

    @property
    def groups(self) -> list[CommutingObservableGroup]:
        return self._groups

    @property
    def lookup(self) -> dict:
        """Dict ``Pauli -> [(group, term)]`` as in the reference (``:250-261``)."""
        out = {}
        for gi, group in enumerate(self._groups):
            for ti, obs in enumerate(group.commuting_observables):
                out.setdefault(obs, []).append((gi, ti))
        return out
