"""The five BASELINE.json configurations as seeded synthetic workloads (shared by bench.py and tests).

A workload fixes, per subsystem, the observables (hence the commuting groups, masks and lookup),
the number of subexperiment tuples C *per rank*, the shots per PUB and the register widths.  Shot
bytes are synthetic: drawn on the device (``torch.Generator`` seeded per rank) when the data is to
be resident in HBM, and mirrored to host ``BitArray`` objects for the end-to-end / CPU-baseline legs.

Sizes follow SURVEY.md section 8d; where BASELINE.json leaves a size open the choice is stated in
``Workload.describe()`` and echoed in bench.py's ``config``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from qiskit_addon_cutting_b200.containers import BitArray, DataBin, PauliList, PrimitiveResult, PubResult, WeightType
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection

_ALIGN = 16


# ------------------------------------------------------------------------------------------------
# Observables of the five configurations
# ------------------------------------------------------------------------------------------------
def heavy_hex_127_edges() -> list[tuple[int, int]]:
    """Coupling map of a 127-qubit heavy-hex lattice: 7 rows (14, 15 x 5, 14) joined by 6 x 4 bridges -> 144 edges."""
    rows, bridges, q = [], [], 0
    for r, n in enumerate([14, 15, 15, 15, 15, 15, 14]):
        rows.append(list(range(q, q + n)))
        q += n
        if r < 6:
            bridges.append(list(range(q, q + 4)))
            q += 4
    assert q == 127
    edges = []
    for row in rows:
        edges += list(zip(row[:-1], row[1:]))
    for r, bridge in enumerate(bridges):
        cols = [0, 4, 8, 12] if r % 2 == 0 else [2, 6, 10, 14]
        for b, c in zip(bridge, cols):
            top = rows[r][min(c, len(rows[r]) - 1)]
            bottom_row = rows[r + 1]
            bottom = bottom_row[min(c, len(bottom_row) - 1)]
            edges += [(top, b), (b, bottom)]
    assert len(edges) == 144
    return edges


def _z_string(n: int, qubits) -> str:
    chars = ["I"] * n
    for qb in qubits:
        chars[n - 1 - qb] = "Z"
    return "".join(chars)


def _restrict(labels: list[str], qubits: list[int], n: int) -> PauliList:
    """Sub-observables on `qubits` (ascending) of n-qubit labels (reference observables_restricted_to_subsystem)."""
    out = []
    for lab in labels:
        out.append("".join(lab[n - 1 - qb] for qb in reversed(qubits)))
    return PauliList(out)


def observables_tutorial() -> dict:
    # docs/tutorials/01_gate_cutting_to_reduce_circuit_width.ipynb cell 8 (partition "AABB")
    return {
        "A": PauliList(["II", "ZI", "ZZ", "XI", "ZZ", "IX"]),
        "B": PauliList(["ZZ", "IZ", "II", "XI", "ZI", "IX"]),
    }


def observables_wire() -> dict:
    # docs/how-tos/how_to_specify_cut_wires.ipynb shape: 3 subsystems (5, 4, 4 qubits), three Z-type observables
    return {
        0: PauliList(["ZIIII", "IIIII", "IIIZZ"]),
        1: PauliList(["IIII", "ZIIZ", "IIII"]),
        2: PauliList(["IIIZ", "IZII", "ZIII"]),
    }


def observables_wide(seed: int, dense: bool) -> dict:
    """2 x 64 qubits, one 100-term observable list.  dense: Z-strings of weight ~32 per half (one group of 100
    terms); local: random X/Y/Z strings of weight 1-4 per half (many small groups)."""
    rng = np.random.default_rng(seed)
    out = {}
    for label in ("A", "B"):
        labels = []
        for _ in range(100):
            chars = ["I"] * 64
            if dense:
                for qb in np.flatnonzero(rng.random(64) < 0.5):
                    chars[qb] = "Z"
            else:
                for qb in rng.choice(64, size=int(rng.integers(1, 5)), replace=False):
                    chars[qb] = "XYZ"[int(rng.integers(3))]
            labels.append("".join(chars))
        out[label] = PauliList(labels)
    return out


def observables_heavy_hex() -> dict:
    """127 single-Z and 144 edge ZZ observables on a heavy-hex lattice cut into qubits 0..62 | 63..126."""
    n = 127
    labels = [_z_string(n, [qb]) for qb in range(n)] + [_z_string(n, e) for e in heavy_hex_127_edges()]
    return {"A": _restrict(labels, list(range(63)), n), "B": _restrict(labels, list(range(63, 127)), n)}


def observables_molecular(seed: int = 505, n_terms: int = 1000) -> dict:
    """1000 Jordan-Wigner-like strings on 64 qubits (X/Y end points joined by Z tails, or ZZ / Z number terms),
    split over 4 subsystems of 16 qubits."""
    rng = np.random.default_rng(seed)
    n = 64
    labels = []
    while len(labels) < n_terms:
        kind = rng.random()
        chars = ["I"] * n
        if kind < 0.15:
            for qb in rng.choice(n, size=int(rng.integers(1, 3)), replace=False):
                chars[qb] = "Z"
        else:
            pairs = 1 if kind < 0.6 else 2
            pts = np.sort(rng.choice(n, size=2 * pairs, replace=False))
            for a, b in pts.reshape(pairs, 2):
                span = min(int(b), int(a) + int(rng.integers(2, 12)))
                chars[a] = "XY"[int(rng.integers(2))]
                chars[span] = "XY"[int(rng.integers(2))]
                for qb in range(a + 1, span):
                    chars[qb] = "Z"
        labels.append("".join(chars))
    return {f"S{j}": _restrict(labels, list(range(16 * j, 16 * j + 16)), n) for j in range(4)}


# ------------------------------------------------------------------------------------------------
# Workload = observables + sizes
# ------------------------------------------------------------------------------------------------
@dataclass
class Workload:
    name: str
    baseline_config: str  # which entry of BASELINE.json "configs" this is
    observables: dict
    n_coeffs: int  # subexperiment tuples per rank
    shots: int
    qpd_bits: int
    kappa: float  # sum |c_i| over the global coefficient list
    seed: int

    def collections(self) -> list:
        if not hasattr(self, "_collections"):
            self._collections = [ObservableCollection(obs) for obs in self.observables.values()]
        return self._collections

    @property
    def n_obs(self) -> int:
        return len(next(iter(self.observables.values())))

    def coefficients(self, rank: int = 0, world: int = 1) -> np.ndarray:
        """Global coefficient list (C * world entries, sum |c| = kappa); rank r owns slice r."""
        return self.coefficients_global(self.n_coeffs * world)

    def coefficients_global(self, n: int) -> np.ndarray:
        """``n`` seeded coefficients with sum |c| = kappa, sorted by weight like the reference's sampled tuples."""
        rng = np.random.default_rng(self.seed + 7)
        counts = rng.multinomial(max(10 * n, 1000), np.ones(n) / n).astype(np.float64)
        counts = -np.sort(-counts)  # generate_cutting_experiments sorts samples by frequency (cutting_experiments.py:135)
        signs = rng.choice([-1.0, 1.0], size=n)
        return signs * counts / counts.sum() * self.kappa

    def describe(self) -> dict:
        colls = self.collections()
        return {
            "workload": self.name,
            "baseline_config": self.baseline_config,
            "subsystems": len(colls),
            "observables": self.n_obs,
            "groups_per_subsystem": [len(c.groups) for c in colls],
            "terms_per_group_max": max(len(g.pauli_bitmasks) for c in colls for g in c.groups),
            "obs_bytes_per_shot": sorted({(g.num_measured_bits + 7) // 8 for c in colls for g in c.groups}),
            "qpd_bytes_per_shot": (self.qpd_bits + 7) // 8,
            "coefficient_tuples_per_gpu": self.n_coeffs,
            "shots_per_subexperiment": self.shots,
            "pubs_per_gpu": sum(self.n_coeffs * len(c.groups) for c in colls),
        }


def get_workload(name: str, scale: float = 1.0) -> Workload:
    """The five configs.  ``scale`` shrinks the tuple count (tests); 1.0 is the benchmark size."""

    def c(n):
        return max(1, int(round(n * scale)))

    if name == "cfg1_tutorial":
        return Workload(name, "configs[0] tutorial gate cutting AABB", observables_tutorial(), 36, 4096, 1, 9.0, 101)
    if name == "cfg2_wire":
        return Workload(name, "configs[1] wire cutting, 3 subsystems, 10^4 shots", observables_wire(), c(512), 10_000, 3, 64.0, 202)
    if name == "cfg3_wide_dense":
        return Workload(name, "configs[2] 2x64 qubits, 10^6 shots, 100-term observable (dense Z strings, one group)",
                        observables_wide(303, True), c(36), 1_000_000, 2, 9.0, 303)
    if name == "cfg3_wide_local":
        return Workload(name, "configs[2] 2x64 qubits, 10^6 shots, 100-term observable (weight 1-4 XYZ, many groups)",
                        observables_wide(304, False), c(36), 1_000_000, 2, 9.0, 304)
    if name == "cfg4_heavy_hex":
        return Workload(name, "configs[3] 127-qubit heavy-hex in 2 halves, 8 cut CNOTs, ~10^5 sampled tuples",
                        observables_heavy_hex(), c(97_000), 8192, 8, 6561.0, 404)
    if name == "cfg4u_heavy_hex_unpartitioned":
        # not a BASELINE.json config: the same 127-qubit observables with the circuit NOT separated (call form 1 of
        # the reference): one subsystem, 16-byte observable rows, one group of 271 terms -> wide-row kernel
        n = 127
        labels = [_z_string(n, [qb]) for qb in range(n)] + [_z_string(n, e) for e in heavy_hex_127_edges()]
        return Workload(name, "extra: configs[3] observables, un-partitioned (16-byte rows)", {"A": PauliList(labels)},
                        c(48_000), 8192, 8, 6561.0, 414)
    if name == "cfg5_molecular":
        w = Workload(name, "configs[4] 1000-term Hamiltonian, 4 subsystems, 10^4 subexperiments x 10^5 shots",
                     observables_molecular(), 1, 100_000, 4, 81.0, 505)
        groups = max(len(col.groups) for col in w.collections())
        w.n_coeffs = c(max(1, 10_000 // groups))
        return w
    raise KeyError(name)


WORKLOADS = ["cfg1_tutorial", "cfg2_wire", "cfg3_wide_dense", "cfg3_wide_local", "cfg4_heavy_hex", "cfg5_molecular",
             "cfg4u_heavy_hex_unpartitioned"]


# ------------------------------------------------------------------------------------------------
# Tables and data
# ------------------------------------------------------------------------------------------------
def shard(w: Workload, rank: int, world: int, scaling: str = "weak"):
    """One rank's share of the global problem: ``(local workload, global coefficient vector, (lo, hi))``.

    weak: every rank holds ``w.n_coeffs`` tuples of a global list of ``world * w.n_coeffs``; strong: the global list has
    ``w.n_coeffs`` tuples and rank r holds the r-th equal-count slice (what ``reconstruct_expectation_values`` with a
    process group assigns to it)."""
    import copy

    from qiskit_addon_cutting_b200.distributed import coefficient_slice

    if scaling == "weak" or world == 1:
        total = w.n_coeffs * world
        lo, hi = rank * w.n_coeffs, (rank + 1) * w.n_coeffs
        local = w
    else:
        total = w.n_coeffs
        lo, hi = coefficient_slice(total, rank, world)
        local = copy.copy(w)
        local.n_coeffs = hi - lo
    return local, w.coefficients_global(total), (lo, hi)


def build_tables(w: Workload, rank: int = 0, world: int = 1, *, coeffs: np.ndarray | None = None):
    """HostTables of one rank's share, built without walking Python PUB objects (vectorised).  ``coeffs``: this
    rank's ``w.n_coeffs`` coefficients (default: slice ``rank`` of the weak-scaling global list)."""
    from qiskit_addon_cutting_b200 import engine
    from qiskit_addon_cutting_b200._lib import PUB_DTYPE

    colls = w.collections()
    builder = engine.TableBuilder(w.n_obs)
    nbq = (w.qpd_bits + 7) // 8
    chunks = []
    offset = 0
    tally = 0
    n_pubs = 0
    for coll in colls:
        G = len(coll.groups)
        li = builder.add_label(coll, n_pubs)
        nb = np.array([(g.num_measured_bits + 7) // 8 for g in coll.groups], dtype=np.int64)
        T = np.array([len(g.pauli_bitmasks) for g in coll.groups], dtype=np.int64)
        var = np.array([builder.group_variant(li, g, int(nb[g])) for g in range(G)], dtype=np.uint32)
        pubs = np.zeros(w.n_coeffs * G, dtype=PUB_DTYPE)
        g_idx = np.tile(np.arange(G), w.n_coeffs)
        obs_sz = (w.shots * nb[g_idx] + _ALIGN - 1) // _ALIGN * _ALIGN
        qpd_sz = np.full(len(pubs), (w.shots * nbq + _ALIGN - 1) // _ALIGN * _ALIGN, dtype=np.int64)
        sizes = np.stack([obs_sz, qpd_sz], axis=1).reshape(-1)
        starts = offset + np.concatenate([[0], np.cumsum(sizes)[:-1]])
        pubs["obs_offset"] = starts[0::2]
        pubs["qpd_offset"] = starts[1::2]
        pubs["tally_offset"] = tally + np.concatenate([[0], np.cumsum(T[g_idx])[:-1]])
        pubs["shots"] = w.shots
        pubs["group"] = var[g_idx]
        pubs["nb_qpd"] = nbq
        offset += int(sizes.sum())
        tally += int(T[g_idx].sum())
        n_pubs += len(pubs)
        chunks.append(pubs)
    if coeffs is None:
        lo = rank * w.n_coeffs
        coeffs = w.coefficients(rank, world)[lo : lo + w.n_coeffs]
    assert len(coeffs) == w.n_coeffs
    return builder.tables(np.concatenate(chunks), coeffs, tally, max(offset, _ALIGN))


def _pub_families(tables):
    """PUB rows in arithmetic progression: yields (first_row, count, row_stride) per (label, group) so that whole families
    of equally shaped PUBs can be addressed as one strided device view."""
    labels = tables.labels
    ends = [int(lab["pub_base"]) for lab in labels[1:]] + [len(tables.pubs)]
    for lab, end in zip(labels, ends):
        G, base = int(lab["num_groups"]), int(lab["pub_base"])
        if G == 0 or end <= base:
            continue
        for g in range(G):
            yield base + g, (end - base) // G, G


def device_data(tables, seed: int, *, biased: bool = True, chunk_bytes: int = 1 << 30):
    """Seeded synthetic shot bytes resident in HBM.

    ``biased`` (default): every bit of every row is Bernoulli(p) with a per-(PUB, bit) bias
    p in {0.25, 0.75}, so single-qubit expectation values are +-0.5 and the reference's sequential fp64 rounding noise
    is exercised; of each PUB's qpd row only a random ~quarter of the bits is ever set (the cut measurements its basis
    elements contain), the other bits are constant 0 like the registers a sampler returns.  Pad bits of the
    observable rows stay random on purpose (masks must ignore them).  ``biased=False``: uniform bytes.
    """
    import torch

    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    n = tables.data_bytes + 128  # QCUT_DATA_SLACK: readable slack after the last matrix

    def rand_bytes(count):
        words = (count + 7) // 8
        return torch.randint(-(2**63), 2**63 - 1, (words,), dtype=torch.int64, device="cuda", generator=gen).view(torch.uint8)[:count]

    data = rand_bytes(n)
    if not biased or len(tables.pubs) == 0:
        return data
    pubs = tables.pubs
    nb_of = tables.groups["nb_obs"][pubs["group"]].astype(np.int64)
    for first, count, stride in _pub_families(tables):
        rows = pubs[first: first + count * stride: stride]
        S, nb, nbq = int(rows["shots"][0]), int(nb_of[first]), int(rows["nb_qpd"][0])
        uniform = count == 1 or (
            np.all(rows["shots"] == S) and np.all(rows["nb_qpd"] == nbq)
            and np.all(np.diff(rows["obs_offset"].astype(np.int64)) == int(rows["obs_offset"][1]) - int(rows["obs_offset"][0]))
            and np.all(np.diff(rows["qpd_offset"].astype(np.int64)) == int(rows["qpd_offset"][1]) - int(rows["qpd_offset"][0])))
        if not uniform or S == 0:
            raise ValueError("device_data(biased=True) expects equally shaped PUBs per (label, group)")
        step = int(rows["obs_offset"][1]) - int(rows["obs_offset"][0]) if count > 1 else S * nb
        for which, width in (("obs_offset", nb), ("qpd_offset", nbq)):
            per = max(1, chunk_bytes // max(1, S * width))
            for c0 in range(0, count, per):
                c1 = min(count, c0 + per)
                view = torch.as_strided(data, (c1 - c0, S, width), (step, width, 1), int(rows[which][c0]))
                r2 = rand_bytes((c1 - c0) * S * width).view(c1 - c0, S, width)
                sel = rand_bytes((c1 - c0) * width).view(c1 - c0, 1, width)
                # bit = (r1 & (r2 | sel)) | (r2 & sel): sel = 0 -> r1 & r2 (p = 1/4), sel = 1 -> r1 | r2 (p = 3/4)
                view.bitwise_and_(r2 | sel)
                view.bitwise_or_(r2.bitwise_and_(sel))
                if which == "qpd_offset":
                    active = rand_bytes((c1 - c0) * width).view(c1 - c0, 1, width) & rand_bytes((c1 - c0) * width).view(c1 - c0, 1, width)
                    view.bitwise_and_(active)
                del r2
    return data


def slice_tables(tables, c_lo: int, c_hi: int, coeffs: np.ndarray | None = None):
    """HostTables of coefficient tuples ``[c_lo, c_hi)`` of every subsystem, addressing the SAME staged buffer
    (strong-scaling legs: rank r of N works on C/N of the resident tuples)."""
    import dataclasses

    labels = tables.labels.copy()
    ends = [int(lab["pub_base"]) for lab in tables.labels[1:]] + [len(tables.pubs)]
    chunks, base, tally = [], 0, 0
    for li, (lab, end) in enumerate(zip(tables.labels, ends)):
        G, pb = int(lab["num_groups"]), int(lab["pub_base"])
        rows = tables.pubs[pb + c_lo * G: min(pb + c_hi * G, end)].copy()
        T = tables.groups["n_terms"][rows["group"]].astype(np.int64)
        rows["tally_offset"] = tally + np.concatenate([[0], np.cumsum(T)[:-1]]) if len(rows) else 0
        tally += int(T.sum())
        labels[li]["pub_base"] = base
        base += len(rows)
        chunks.append(rows)
    sub = tables.coeffs[c_lo:c_hi] if coeffs is None else np.ascontiguousarray(coeffs, dtype=np.float64)
    return dataclasses.replace(tables, pubs=np.concatenate(chunks), labels=labels, coeffs=sub, n_tallies=tally)


def host_results(w: Workload, tables, fetch, *, total_coeffs: int | None = None, first_coeff: int = 0,
                 max_coeffs: int | None = None, max_shots: int | None = None) -> dict:
    """Per-label ``PrimitiveResult`` holding exactly the bytes the kernels saw, every ``BitArray`` in an allocation of
    its own (as a sampler returns them -- 2 x #PUB separate NumPy arrays, not views of one buffer).

    ``fetch(lo, hi)`` returns bytes ``[lo, hi)`` of the staged layout as a host array.  With ``total_coeffs`` the
    result describes a GLOBAL problem of that many tuples of which this process holds tuples ``first_coeff ..
    first_coeff + n_coeffs - 1``; the other entries are absent (``distributed.ShardedResult``).  ``max_coeffs`` /
    ``max_shots`` cut a sub-problem out (first tuples, first shots of every PUB) for the CPU legs.
    """
    colls = w.collections()
    out = {}
    nbq = (w.qpd_bits + 7) // 8
    ends = [int(lab["pub_base"]) for lab in tables.labels[1:]] + [len(tables.pubs)]
    for label, coll, lab, end in zip(w.observables, colls, tables.labels, ends):
        base = int(lab["pub_base"])
        G = len(coll.groups)
        if max_coeffs is not None:
            end = min(end, base + max_coeffs * G)
        pubs = []
        block = 4096
        for b0 in range(base, end, block):
            b1 = min(end, b0 + block)
            lo = int(tables.pubs["obs_offset"][b0])
            last = tables.pubs[b1 - 1]
            hi = int(last["qpd_offset"]) + int(last["shots"]) * nbq
            chunk = fetch(lo, hi)
            for p in range(b0, b1):
                pub = tables.pubs[p]
                nb = int(tables.groups[pub["group"]]["nb_obs"])
                S = int(pub["shots"]) if max_shots is None else min(int(pub["shots"]), max_shots)
                oo, qo = int(pub["obs_offset"]) - lo, int(pub["qpd_offset"]) - lo
                o = chunk[oo: oo + S * nb].reshape(S, nb).copy()
                q = chunk[qo: qo + S * nbq].reshape(S, nbq).copy()
                pubs.append(PubResult(DataBin(observable_measurements=BitArray(o, nb * 8), qpd_measurements=BitArray(q, nbq * 8))))
        if total_coeffs is not None and total_coeffs != len(pubs) // max(G, 1):
            from qiskit_addon_cutting_b200.distributed import ShardedResult

            out[label] = ShardedResult(total_coeffs * G, first_coeff * G, pubs)
        else:
            out[label] = PrimitiveResult(pubs)
    return out


def device_fetch(data):
    """``fetch`` for :func:`host_results` reading from the resident device buffer."""
    return lambda lo, hi: data[lo:hi].cpu().numpy()


def coefficient_pairs(coeffs: np.ndarray) -> list:
    return [(float(c), WeightType.SAMPLED) for c in coeffs]
