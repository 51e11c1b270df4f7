"""Qiskit-free stand-ins for the objects that cross the reconstruction boundary.

The reference's hot path (``qiskit_addon_cutting/cutting_reconstruction.py:31-170``)
touches Qiskit objects only through a narrow protocol:

* ``PauliList`` / ``Pauli``: ``len()``, iteration, ``.phase``, ``.z`` / ``.x``
  symplectic arrays, ``PauliList.unique()`` and
  ``PauliList.group_commuting(qubit_wise=True)``
  (``utils/observable_grouping.py:176,186``);
* ``PrimitiveResult``: ``len(result)`` and
  ``result[idx].data.<register>.array`` (``cutting_reconstruction.py:152-155``);
* ``SamplerResult``: ``.quasi_dists`` (``cutting_reconstruction.py:122-124,145``).

Qiskit is not installable in the build image (no wheels, no network), so this
module provides minimal containers implementing exactly that protocol.  Real
Qiskit objects satisfy the same protocol and are accepted unchanged by
:func:`qiskit_addon_cutting_b200.reconstruct_expectation_values`.

Conventions restated from Qiskit (``SURVEY.md`` appendix E): Pauli labels are
little-endian (rightmost character is qubit 0); ``phase`` is the exponent of
``-i``; ``BitArray.array`` is ``uint8`` of shape ``(num_shots, ceil(num_bits/8))``
and big-endian along the last axis.
"""

from __future__ import annotations

from collections import defaultdict
from collections.abc import Iterable, Sequence
from enum import Enum

import numpy as np

__all__ = [
    "WeightType",
    "Pauli",
    "PauliList",
    "SparsePauliOp",
    "BitArray",
    "DataBin",
    "PubResult",
    "PrimitiveResult",
    "QuasiDistribution",
    "SamplerResult",
]


class WeightType(Enum):
    """Mirror of ``qiskit_addon_cutting/qpd/weights.py:37-44`` (ignored by reconstruction)."""

    EXACT = 1
    SAMPLED = 2


_PHASE_PREFIX = {"": 0, "-i": 1, "-": 2, "i": 3, "+": 0, "+i": 3}
_ZX_OF_CHAR = {"I": (False, False), "X": (False, True), "Y": (True, True), "Z": (True, False)}
_CHAR_OF_ZX = {v: k for k, v in _ZX_OF_CHAR.items()}


def _split_label(label: str) -> tuple[int, str]:
    body = label.lstrip("+-i")
    prefix = label[: len(label) - len(body)]
    if prefix not in _PHASE_PREFIX or any(c not in _ZX_OF_CHAR for c in body):
        raise ValueError(f"Pauli string label '{label}' is not valid.")
    return _PHASE_PREFIX[prefix], body


class Pauli:
    """An n-qubit Pauli string with a phase ``(-i)**phase``."""

    __slots__ = ("z", "x", "phase")

    def __init__(self, data):
        if isinstance(data, Pauli):
            self.z, self.x, self.phase = data.z.copy(), data.x.copy(), data.phase
        elif isinstance(data, str):
            self.phase, body = _split_label(data)
            zx = [_ZX_OF_CHAR[c] for c in reversed(body)]
            self.z = np.array([p[0] for p in zx], dtype=bool)
            self.x = np.array([p[1] for p in zx], dtype=bool)
        elif isinstance(data, tuple) and len(data) in (2, 3):
            self.z = np.atleast_1d(np.asarray(data[0])).astype(bool)
            self.x = np.atleast_1d(np.asarray(data[1])).astype(bool)
            self.phase = int(data[2]) % 4 if len(data) == 3 else 0
            if self.z.shape != self.x.shape or self.z.ndim != 1:
                raise ValueError("z and x must be 1D arrays of equal length.")
        else:
            raise ValueError("Invalid input data for Pauli.")

    @property
    def num_qubits(self) -> int:
        return self.z.shape[0]

    def __len__(self) -> int:
        return self.z.shape[0]

    def __getitem__(self, qubits) -> "Pauli":
        # Slicing drops the phase, as in Qiskit.
        if isinstance(qubits, (int, np.integer)):
            if not -len(self) <= qubits < len(self):
                raise IndexError("qubit index out of range")
            qubits = [int(qubits)]
        return Pauli((self.z[qubits], self.x[qubits]))

    def __setitem__(self, qubit, value) -> None:
        value = value if isinstance(value, Pauli) else Pauli(value)
        self.z[qubit] = value.z if np.ndim(self.z[qubit]) else value.z[0]
        self.x[qubit] = value.x if np.ndim(self.x[qubit]) else value.x[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def to_label(self) -> str:
        body = "".join(_CHAR_OF_ZX[(bool(z), bool(x))] for z, x in zip(self.z[::-1], self.x[::-1]))
        return {0: "", 1: "-i", 2: "-", 3: "i"}[self.phase] + body

    def __eq__(self, other) -> bool:
        if not isinstance(other, Pauli):
            return NotImplemented
        return (
            self.phase == other.phase
            and self.z.shape == other.z.shape
            and bool(np.all(self.z == other.z))
            and bool(np.all(self.x == other.x))
        )

    def __hash__(self) -> int:
        return hash(self.to_label())

    def __repr__(self) -> str:
        return f"Pauli('{self.to_label()}')"


class PauliList:
    """A list of equal-width Pauli strings stored as symplectic ``z`` / ``x`` matrices."""

    def __init__(self, data):
        if isinstance(data, PauliList):
            self.z, self.x, self.phase = data.z.copy(), data.x.copy(), data.phase.copy()
            return
        if isinstance(data, (str, Pauli)):
            data = [data]
        paulis = [p if isinstance(p, Pauli) else Pauli(p) for p in data]
        if not paulis:
            raise ValueError("Input Pauli list is empty.")
        n = len(paulis[0])
        if any(len(p) != n for p in paulis):
            raise ValueError("All Paulis must have the same number of qubits.")
        self.z = np.array([p.z for p in paulis], dtype=bool).reshape(len(paulis), n)
        self.x = np.array([p.x for p in paulis], dtype=bool).reshape(len(paulis), n)
        self.phase = np.array([p.phase for p in paulis], dtype=np.int64)

    @classmethod
    def from_symplectic(cls, z, x, phase=0) -> "PauliList":
        self = cls.__new__(cls)
        self.z = np.atleast_2d(np.asarray(z)).astype(bool)
        self.x = np.atleast_2d(np.asarray(x)).astype(bool)
        if self.z.shape != self.x.shape:
            raise ValueError("z and x must have the same shape.")
        self.phase = np.broadcast_to(np.asarray(phase, dtype=np.int64) % 4, (self.z.shape[0],)).copy()
        return self

    @property
    def num_qubits(self) -> int:
        return self.z.shape[1]

    @property
    def size(self) -> int:
        return self.z.shape[0]

    def __len__(self) -> int:
        return self.z.shape[0]

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            return Pauli((self.z[index], self.x[index], int(self.phase[index])))
        index = np.asarray(index) if not isinstance(index, slice) else index
        return PauliList.from_symplectic(self.z[index], self.x[index], self.phase[index])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def copy(self) -> "PauliList":
        return PauliList(self)

    def to_labels(self) -> list[str]:
        return [p.to_label() for p in self]

    def __repr__(self) -> str:
        return f"PauliList({self.to_labels()})"

    def __eq__(self, other):
        if not isinstance(other, PauliList):
            return NotImplemented
        return (
            self.z.shape == other.z.shape
            and bool(np.all(self.z == other.z))
            and bool(np.all(self.x == other.x))
            and bool(np.all(self.phase == other.phase))
        )

    __hash__ = None  # mutable container

    def unique(self) -> "PauliList":
        """First occurrences, in original order (Qiskit ``PauliList.unique`` behaviour)."""
        cols = [self.z, self.x]
        if np.any(self.phase != self.phase[0]):
            cols.append(self.phase.reshape(-1, 1))
        _, index = np.unique(np.hstack(cols).astype(np.int8), return_index=True, axis=0)
        return self[np.sort(index)]

    def noncommutation_edges(self, qubit_wise: bool = True) -> list[tuple[int, int]]:
        """Upper-triangle pairs that do not commute (qubit-wise: differ on a shared non-I qubit)."""
        code = self.z.astype(np.int8) + 2 * self.x.astype(np.int8)
        a, b = code[:, None, :], code[None, :, :]
        anti = (a * b) * (a - b)
        if qubit_wise:
            adjacency = np.logical_or.reduce(anti != 0, axis=2)
        else:
            adjacency = np.logical_xor.reduce(anti != 0, axis=2)
        rows, cols = np.where(np.triu(adjacency, k=1))
        return list(zip(rows.tolist(), cols.tolist()))

    def group_commuting(self, qubit_wise: bool = False) -> list["PauliList"]:
        """Greedy colouring of the non-commutation graph.

        Restates Qiskit's ``PauliList.group_commuting`` over
        ``rustworkx.graph_greedy_color``: nodes are visited by descending
        degree (stable in index), each takes the smallest colour unused by its
        neighbours, and groups are emitted in order of first colour appearance
        in that visit order.  Neither Qiskit nor rustworkx is vendored in the
        reference, so this ordering is *parity unpinned* (``SURVEY.md`` §8c).
        """
        n = len(self)
        neighbours: list[set[int]] = [set() for _ in range(n)]
        for i, j in self.noncommutation_edges(qubit_wise):
            neighbours[i].add(j)
            neighbours[j].add(i)
        order = sorted(range(n), key=lambda k: -len(neighbours[k]))
        colour: dict[int, int] = {}
        for node in order:
            used = {colour[m] for m in neighbours[node] if m in colour}
            c = 0
            while c in used:
                c += 1
            colour[node] = c
        groups: dict[int, list[int]] = defaultdict(list)
        for node, c in colour.items():
            groups[c].append(node)
        return [self[g] for g in groups.values()]


_PHASE_FACTOR = np.array([1, -1j, -1, 1j], dtype=np.complex128)  # (-i)**phase


class SparsePauliOp:
    """Linear combination of phase-free Paulis: ``sum_j coeffs[j] * paulis[j]`` (Qiskit ``SparsePauliOp`` protocol).

    Only what the term-recombination step touches (reference ``utils/observable_terms.py:22-90``):
    ``.paulis`` (phases absorbed into ``.coeffs``, as Qiskit does), ``.coeffs`` (complex128),
    ``.num_qubits``, ``len()`` and iteration over single-term operators.
    """

    def __init__(self, data, coeffs=None):
        if isinstance(data, SparsePauliOp):
            self.paulis = PauliList(data.paulis)
            self.coeffs = data.coeffs.copy() if coeffs is None else np.asarray(coeffs, dtype=np.complex128).copy()
            return
        pl = data if isinstance(data, PauliList) else PauliList(data)
        c = np.ones(len(pl), dtype=np.complex128) if coeffs is None else np.asarray(coeffs, dtype=np.complex128)
        if c.shape != (len(pl),):
            raise ValueError("coeff vector is incorrect shape for number of Paulis")
        self.coeffs = _PHASE_FACTOR[pl.phase] * c
        self.paulis = PauliList.from_symplectic(pl.z, pl.x, 0)

    @property
    def num_qubits(self) -> int:
        return self.paulis.num_qubits

    @property
    def size(self) -> int:
        return len(self.paulis)

    def __len__(self) -> int:
        return len(self.paulis)

    def __getitem__(self, index) -> "SparsePauliOp":
        if isinstance(index, (int, np.integer)):
            index = [int(index)]
        return SparsePauliOp(self.paulis[index], self.coeffs[index])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self) -> str:
        return f"SparsePauliOp({self.paulis.to_labels()}, coeffs={self.coeffs.tolist()})"


class BitArray:
    """Packed shot data of one classical register (``qiskit.primitives.BitArray`` layout)."""

    def __init__(self, array: np.ndarray, num_bits: int):
        array = np.asarray(array)
        if array.dtype != np.uint8:
            raise TypeError(f"Input array must have dtype uint8, not {array.dtype}.")
        if array.ndim < 2:
            raise ValueError("The input array must have at least two axes.")
        if array.shape[-1] != (num_bits + 7) // 8:
            raise ValueError("The input array is expected to have ceil(num_bits/8) bytes per shot.")
        self._array = array
        self._num_bits = int(num_bits)

    @property
    def array(self) -> np.ndarray:
        return self._array

    @property
    def num_bits(self) -> int:
        return self._num_bits

    @property
    def num_shots(self) -> int:
        return self._array.shape[-2]

    @property
    def shape(self) -> tuple[int, ...]:
        return self._array.shape[:-2]

    @classmethod
    def from_ints(cls, values: Iterable[int], num_bits: int) -> "BitArray":
        nb = (num_bits + 7) // 8
        rows = [int(v).to_bytes(nb, "big") for v in values]
        return cls(np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), nb), num_bits)


class DataBin:
    """Namespace of per-register data of one PUB."""

    def __init__(self, **fields):
        self.__dict__.update(fields)

    def __repr__(self) -> str:
        return f"DataBin({', '.join(self.__dict__)})"


class PubResult:
    def __init__(self, data: DataBin, metadata: dict | None = None):
        self.data = data
        self.metadata = metadata or {}


class PrimitiveResult:
    """SamplerV2 result container: a sequence of :class:`PubResult`."""

    def __init__(self, pub_results: Iterable[PubResult], metadata: dict | None = None):
        self._pub_results = list(pub_results)
        self.metadata = metadata or {}

    def __len__(self) -> int:
        return len(self._pub_results)

    def __getitem__(self, index):
        return self._pub_results[index]

    def __iter__(self):
        return iter(self._pub_results)


class QuasiDistribution(dict):
    """Outcome -> quasi-probability; string keys are converted to ints as Qiskit does."""

    def __init__(self, data: dict | None = None, shots: int | None = None):
        self.shots = shots
        converted = {}
        for key, value in (data or {}).items():
            if isinstance(key, str):
                if key.startswith("0x"):
                    key = int(key, 16)
                elif key.startswith("0b"):
                    key = int(key, 2)
                else:
                    key = int(key.replace(" ", ""), 2)
            converted[int(key)] = value
        super().__init__(converted)


class SamplerResult:
    """SamplerV1 result container."""

    def __init__(self, quasi_dists: Sequence[dict], metadata: Sequence[dict] | None = None):
        self.quasi_dists = list(quasi_dists)
        self.metadata = list(metadata) if metadata is not None else [{} for _ in self.quasi_dists]
