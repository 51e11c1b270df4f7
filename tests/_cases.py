"""Seeded synthetic inputs for the reconstruction path (shared by tests, golden generation, smoke).

Every case is a triple ``(results, coefficients, observables)`` in the exact
shape ``reconstruct_expectation_values`` takes: per label one
``PrimitiveResult`` whose PUB ``i * G + g`` carries the ``observable_measurements``
(width ``max(1, len(pauli_indices_g))`` bits) and ``qpd_measurements`` registers
(reference ``cutting_experiments.py:140-157,268-270``; ``qpd/decompose.py:147-150``).
"""

from __future__ import annotations

import numpy as np

from qiskit_addon_cutting_b200.containers import (
    BitArray,
    DataBin,
    PauliList,
    PrimitiveResult,
    PubResult,
    WeightType,
)
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection


def random_bits(rng: np.random.Generator, shots: int, num_bits: int) -> np.ndarray:
    """``(shots, ceil(num_bits/8))`` uint8, big-endian rows, per-bit Bernoulli bias in [0.2, 0.8], pad bits zero."""
    nb = (num_bits + 7) // 8
    bias = rng.uniform(0.2, 0.8, size=num_bits)
    bits = rng.random((shots, num_bits)) < bias[None, :]
    # bit b of the register is bit (b % 8) of byte (nb - 1 - b // 8)
    padded = np.zeros((shots, nb * 8), dtype=bool)
    padded[:, :num_bits] = bits
    little = np.packbits(padded, axis=1, bitorder="little")  # byte 0 holds bits 0..7
    return np.ascontiguousarray(little[:, ::-1])


def make_results(
    observables: dict,
    num_coeffs: int,
    shots,
    qpd_bits,
    seed: int,
) -> dict:
    """PrimitiveResults consistent with our grouping of ``observables``.

    ``shots`` and ``qpd_bits`` may be ints or callables ``(label, i, g) -> int``.
    """
    rng = np.random.default_rng(seed)
    results = {}
    for label, obs in observables.items():
        groups = ObservableCollection(obs).groups
        pubs = []
        for i in range(num_coeffs):
            for g, cog in enumerate(groups):
                s = shots(label, i, g) if callable(shots) else shots
                qb = qpd_bits(label, i, g) if callable(qpd_bits) else qpd_bits
                ob = cog.num_measured_bits
                pubs.append(
                    PubResult(
                        DataBin(
                            observable_measurements=BitArray(random_bits(rng, s, ob), ob),
                            qpd_measurements=BitArray(random_bits(rng, s, qb), qb),
                        )
                    )
                )
        results[label] = PrimitiveResult(pubs)
    return results


def random_coefficients(rng: np.random.Generator, n: int, kappa: float = 9.0) -> list:
    signs = rng.choice([-1.0, 1.0], size=n)
    weights = rng.dirichlet(np.ones(n)) * kappa
    kinds = [WeightType.EXACT, WeightType.SAMPLED]
    return [(float(s * w), kinds[int(rng.integers(2))]) for s, w in zip(signs, weights)]


def random_paulis(rng, num: int, n: int, weight_lo: int, weight_hi: int, alphabet="XYZ") -> PauliList:
    labels = []
    for _ in range(num):
        w = int(rng.integers(weight_lo, weight_hi + 1))
        chars = ["I"] * n
        for q in rng.choice(n, size=min(w, n), replace=False):
            chars[q] = alphabet[int(rng.integers(len(alphabet)))]
        labels.append("".join(chars))
    return PauliList(labels)


# ---------------------------------------------------------------------------------------
# Named small cases: miniature versions of BASELINE.json's five configs plus edge cases.
# ---------------------------------------------------------------------------------------
def case_tutorial(seed=101, shots=64):
    """cfg1 shape: 'AABB' partition, 6 observables (docs/tutorials/01 cell 8), 36 tuples of +-0.25."""
    observables = {
        "A": PauliList(["II", "ZI", "ZZ", "XI", "ZZ", "IX"]),
        "B": PauliList(["ZZ", "IZ", "II", "XI", "ZI", "IX"]),
    }
    coeffs = []
    for a in range(6):
        for b in range(6):
            neg = ((a in (3, 5)) + (b in (3, 5))) % 2
            coeffs.append((-0.25 if neg else 0.25, WeightType.EXACT))
    results = make_results(observables, 36, shots, 1, seed)
    return results, coeffs, observables


def case_wire(seed=202, shots=100):
    """cfg2 shape: 3 subsystems, Z-type observables, 1-3 bit qpd registers, non power-of-two shots."""
    observables = {
        0: PauliList(["ZIIII", "IIIII", "IIIZZ"]),
        1: PauliList(["IIII", "ZIIZ", "IIII"]),
        2: PauliList(["IIIZ", "IZII", "ZIII"]),
    }
    rng = np.random.default_rng(seed)
    coeffs = [(float(c), WeightType.EXACT) for c in rng.choice([-0.125, 0.125], size=16)]
    results = make_results(observables, 16, shots, lambda label, i, g: 1 + (i + label) % 3, seed + 1)
    return results, coeffs, observables


def case_wide(seed=303, shots=257, dense=True):
    """cfg3 shape: 2 subsystems x 64 qubits, 8-byte observable register."""
    rng = np.random.default_rng(seed)
    if dense:
        observables = {
            "A": random_paulis(rng, 12, 64, 24, 40, "Z"),
            "B": random_paulis(rng, 12, 64, 24, 40, "Z"),
        }
    else:
        observables = {
            "A": random_paulis(rng, 12, 64, 1, 4),
            "B": random_paulis(rng, 12, 64, 1, 4),
        }
    coeffs = random_coefficients(rng, 4)
    results = make_results(observables, 4, shots, 2, seed + 1)
    return results, coeffs, observables


def case_heavy_hex(seed=404, shots=128):
    """cfg4 shape: 63 + 64 qubits, single-Z and ZZ terms, 8 cut gates -> up to 8 qpd bits."""
    rng = np.random.default_rng(seed)

    def zz(n, pairs):
        out = []
        for a, b in pairs:
            chars = ["I"] * n
            chars[n - 1 - a] = "Z"
            if b is not None:
                chars[n - 1 - b] = "Z"
            out.append("".join(chars))
        return out

    pa = [(int(q), None) for q in rng.choice(63, 10, replace=False)] + [
        (int(q), int(q) + 1) for q in rng.choice(62, 10, replace=False)
    ]
    pb = [(int(q), None) for q in rng.choice(64, 10, replace=False)] + [
        (int(q), int(q) + 1) for q in rng.choice(63, 10, replace=False)
    ]
    observables = {"A": PauliList(zz(63, pa)), "B": PauliList(zz(64, pb))}
    coeffs = random_coefficients(rng, 6, kappa=6561.0)
    results = make_results(observables, 6, shots, 8, seed + 1)
    return results, coeffs, observables


def case_molecular(seed=505, shots=96):
    """cfg5 shape: 4 subsystems x 16 qubits, many mixed Pauli terms, several groups per subsystem."""
    rng = np.random.default_rng(seed)
    observables = {f"S{j}": random_paulis(rng, 40, 16, 0, 6) for j in range(4)}
    coeffs = random_coefficients(rng, 5)
    results = make_results(observables, 5, shots, 3, seed + 1)
    return results, coeffs, observables


def case_odd_widths(seed=606, shots=77):
    """Ragged edge cases: 3-, 2- and 5-byte observable rows, 2-byte qpd rows, identity-only group, duplicates."""
    rng = np.random.default_rng(seed)
    a = random_paulis(rng, 7, 20, 1, 20, "Z")
    b_labels = random_paulis(rng, 5, 9, 1, 9, "XZ").to_labels() + ["I" * 9, "I" * 9]
    c = random_paulis(rng, 7, 33, 1, 33, "Z")
    observables = {"a": a, ("b", 1): PauliList(b_labels), 3: c}
    coeffs = random_coefficients(rng, 3)
    results = make_results(
        observables, 3, lambda label, i, g: shots + 3 * i + g, lambda label, i, g: 11 if label == "a" else 1, seed + 1
    )
    return results, coeffs, observables


def case_identity_only(seed=707, shots=50):
    """Every sub-observable is the identity on subsystem 'B' (reference issue #422 regression shape)."""
    observables = {"A": PauliList(["ZZ", "XX"]), "B": PauliList(["II", "II"])}
    rng = np.random.default_rng(seed)
    coeffs = random_coefficients(rng, 4)
    results = make_results(observables, 4, shots, 2, seed + 1)
    return results, coeffs, observables


def case_unpartitioned(seed=808, shots=130):
    """Call form 1: a bare PauliList and a single PrimitiveResult (label 'A' implied)."""
    rng = np.random.default_rng(seed)
    observables = random_paulis(rng, 9, 6, 0, 6)
    coeffs = random_coefficients(rng, 7)
    results = make_results({"A": observables}, 7, shots, 4, seed + 1)["A"]
    return results, coeffs, observables


def case_pow2(seed=909, shots=256):
    """Power-of-two shots: the reference's sequential fp64 accumulation is exact here."""
    rng = np.random.default_rng(seed)
    observables = {"L": random_paulis(rng, 10, 12, 1, 5), "R": random_paulis(rng, 10, 10, 1, 5)}
    coeffs = [(float(c), WeightType.EXACT) for c in rng.choice([-0.5, 0.5, 0.25], size=6)]
    results = make_results(observables, 6, shots, 2, seed + 1)
    return results, coeffs, observables


SMALL_CASES = {
    "tutorial": case_tutorial,
    "wire": case_wire,
    "wide_dense": lambda: case_wide(dense=True),
    "wide_local": lambda: case_wide(seed=304, dense=False),
    "heavy_hex": case_heavy_hex,
    "molecular": case_molecular,
    "odd_widths": case_odd_widths,
    "identity_only": case_identity_only,
    "unpartitioned": case_unpartitioned,
    "pow2": case_pow2,
}


# ---- term recombination (reference utils/observable_terms.py) ------------------------------------------
def terms_case(seed: int, complex_values: bool, n_qubits: int = 5, pool: int = 40, n_observables: int = 14):
    """Seeded observables (SparsePauliOps with complex coefficients, zero coefficients, phased labels, bare
    Paulis) and the expectation value of every unique term.  Returns (observables, {Pauli: value})."""
    from qiskit_addon_cutting_b200.containers import Pauli, SparsePauliOp

    rng = np.random.default_rng(seed)
    labels = sorted({"".join(rng.choice(list("IXYZ"), size=n_qubits)) for _ in range(pool)})
    observables = []
    for k in range(n_observables):
        if k % 5 == 4:
            observables.append(Pauli(["", "-", "i", "-i"][int(rng.integers(4))] + labels[int(rng.integers(len(labels)))]))
            continue
        n_terms = int(rng.integers(1, 70 if k % 3 == 0 else 9))  # some rows longer than one warp pass
        chosen = [labels[int(i)] for i in rng.integers(len(labels), size=n_terms)]  # repeats allowed
        chosen = [["", "-", "i"][int(rng.integers(3))] + c for c in chosen]
        coeffs = rng.normal(size=n_terms) + 1j * rng.normal(size=n_terms) * (rng.random(n_terms) < 0.5)
        coeffs[rng.random(n_terms) < 0.15] = 0
        observables.append(SparsePauliOp(chosen, coeffs))
    values = {}
    for lab in labels:
        v = float(rng.uniform(-1, 1))
        values[Pauli(lab)] = complex(v, float(rng.uniform(-1, 1))) if complex_values else v
    return observables, values


TERMS_CASES = {
    "real": lambda: terms_case(1201, False),
    "complex": lambda: terms_case(1202, True),
    "single_qubit": lambda: terms_case(1203, False, n_qubits=1, pool=8, n_observables=6),
}


# ---- SamplerV1 views of SamplerV2 results (for inputs that mix the two, reference cutting_reconstruction.py:143,150) ----
def to_sampler_v1(result, observables):
    """``SamplerResult`` whose quasi-distributions are the outcome frequencies of a ``PrimitiveResult``'s PUBs: the
    outcome integer is ``(qpd << num_measured_bits) | obs`` (reference ``_process_outcome``, :173-193)."""
    from qiskit_addon_cutting_b200.containers import QuasiDistribution, SamplerResult

    groups = ObservableCollection(observables).groups
    dists = []
    for idx in range(len(result)):
        data = result[idx].data
        nmb = groups[idx % len(groups)].num_measured_bits
        obs, qpd = data.observable_measurements.array, data.qpd_measurements.array
        counts: dict = {}
        for j in range(qpd.shape[0]):
            o = int.from_bytes(obs[j].tobytes(), "big") & ((1 << nmb) - 1)
            key = (int.from_bytes(qpd[j].tobytes(), "big") << nmb) | o
            counts[key] = counts.get(key, 0) + 1
        shots = max(qpd.shape[0], 1)
        dists.append(QuasiDistribution({k: v / shots for k, v in counts.items()}))
    return SamplerResult(dists)


def case_mixed_samplers(seed=1010, shots=60):
    """Three subsystems: SamplerV2, SamplerV1, SamplerV2 -- the reference handles each subsystem by its own type."""
    rng = np.random.default_rng(seed)
    observables = {"A": random_paulis(rng, 8, 7, 0, 4), "B": random_paulis(rng, 8, 5, 0, 4), "C": random_paulis(rng, 8, 9, 1, 5, "Z")}
    coeffs = random_coefficients(rng, 5)
    results = make_results(observables, 5, lambda label, i, g: shots + i, 2, seed + 1)
    results["B"] = to_sampler_v1(results["B"], observables["B"])
    return results, coeffs, observables
