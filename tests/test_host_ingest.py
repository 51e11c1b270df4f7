"""Host-side ingestion (no GPU): the C-API result walk, table building and gather reproduce the staged layout."""

import numpy as np
import pytest

import _cases
from oracle import c_oracle, oracle
from qiskit_addon_cutting_b200 import engine
from qiskit_addon_cutting_b200.containers import BitArray, DataBin, PauliList, PrimitiveResult, PubResult, WeightType
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection


def _tables(results, coefficients, observables):
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)
    return engine.build_host_tables([results[label] for label in labels], colls, coeffs, 0, len(coeffs),
                                    len(observables[labels[0]]))


@pytest.mark.parametrize("name", ["molecular", "odd_widths", "heavy_hex"])
def test_staged_layout_feeds_the_c_oracle(lib, name):
    """tables + gathered bytes are self-consistent: the C oracle run on them gives the oracle's tallies."""
    results, coefficients, observables = _cases.SMALL_CASES[name]()
    tables, arrays = _tables(results, coefficients, observables)
    data = np.zeros(tables.data_bytes, dtype=np.uint8)
    engine.gather_arrays(tables, arrays, data.ctypes.data, 3)
    _, tally = c_oracle.pubs(data, tables.pubs, tables.groups, tables.masks, tables.n_tallies, want_seq=False, n_threads=2)
    want = np.concatenate([t for label in observables for t in oracle.all_tallies(results, coefficients, observables)[label]])
    assert np.array_equal(tally, want)
    assert tables.single_slot and tables.work == int((tables.pubs["shots"] * tables.groups["n_terms"][tables.pubs["group"]]).sum())


def test_non_conforming_arrays_take_the_slow_path(lib):
    """Fortran-ordered / strided / wider-than-needed arrays are copied to C-contiguous uint8 rows first."""
    rng = np.random.default_rng(3)
    observables = {"A": PauliList(["ZIZ", "IZZ", "ZZZ"])}
    base = rng.integers(0, 8, size=(40, 1), dtype=np.uint8)
    qpd = rng.integers(0, 4, size=(33, 1), dtype=np.uint8)

    class Reg:  # a register whose .array is not a plain C-contiguous matrix
        def __init__(self, array):
            self.array = array

    strided = np.asfortranarray(np.repeat(base, 2, axis=1))[:, :1]  # not C-contiguous, 40 rows (> 33 qpd rows)
    results = {"A": PrimitiveResult([PubResult(DataBin(observable_measurements=Reg(strided), qpd_measurements=Reg(qpd)))])}
    coefficients = [(1.0, WeightType.EXACT)]
    tables, arrays = _tables(results, coefficients, observables)
    assert int(tables.pubs["shots"][0]) == 33  # shots come from the qpd register (reference :155)
    data = np.zeros(tables.data_bytes, dtype=np.uint8)
    engine.gather_arrays(tables, arrays, data.ctypes.data, 1)
    _, tally = c_oracle.pubs(data, tables.pubs, tables.groups, tables.masks, tables.n_tallies, want_seq=False)
    ref = {"A": PrimitiveResult([PubResult(DataBin(observable_measurements=BitArray(np.ascontiguousarray(base[:33]), 3),
                                                   qpd_measurements=BitArray(qpd, 2)))])}
    assert np.array_equal(tally, oracle.all_tallies(ref, coefficients, observables)["A"][0])
    # fewer observable rows than qpd rows: the reference raises IndexError (obs_array[j])
    short = {"A": PrimitiveResult([PubResult(DataBin(observable_measurements=Reg(base[:10]), qpd_measurements=Reg(qpd)))])}
    with pytest.raises(IndexError):
        _tables(short, coefficients, observables)


@pytest.mark.parametrize("name", ["molecular", "odd_widths"])
def test_blocked_walk_gives_the_same_tables(lib, name):
    """Walking the PUBs in small blocks changes nothing in the tables or the staged layout."""
    results, coefficients, observables = _cases.SMALL_CASES[name]()
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)
    args = ([results[label] for label in labels], colls, coeffs, 0, len(coeffs), len(observables[labels[0]]))
    whole, arrays_w = engine.build_host_tables(*args)
    for block in (1, 3, 7):
        tables, arrays = engine.build_host_tables(*args, block_pubs=block)
        for f in ("groups", "masks", "pubs", "labels", "lookup_start", "lookup", "coeffs"):
            assert np.array_equal(getattr(tables, f), getattr(whole, f)), (block, f)
        assert (tables.n_tallies, tables.data_bytes) == (whole.n_tallies, whole.data_bytes)
        assert np.array_equal(arrays.dst, arrays_w.dst) and np.array_equal(arrays.nbytes, arrays_w.nbytes)


def test_block_sink_offsets_are_relative_to_the_lowest_block(lib):
    """With a sink every block lands in its own allocation; PUB offsets are relative to the lowest address."""
    results, coefficients, observables = _cases.SMALL_CASES["molecular"]()
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)

    class Sink:
        def __init__(self):
            self.blocks, self.base = [], 0

        def __call__(self, src, nbytes, dst, block_bytes):
            # over-allocate by a varying amount so that the blocks are neither adjacent nor ordered by size
            block = np.zeros(block_bytes + 64 * (1 + len(self.blocks) % 3), dtype=np.uint8)
            self.blocks.append(block)
            assert lib.qcut_gather_host(src.ctypes.data, nbytes.ctypes.data, dst.ctypes.data, len(src), block.ctypes.data, 1) == 0
            return block.ctypes.data

        def rebase(self, addresses):
            self.base = min(addresses)
            return [a - self.base for a in addresses]

    sink = Sink()
    tables, _ = engine.build_host_tables([results[label] for label in labels], colls, coeffs, 0, len(coeffs),
                                         len(observables[labels[0]]), sink=sink, block_pubs=5)
    assert len(sink.blocks) > len(labels)
    # read every PUB back through base + offset and compare with the caller's arrays
    import ctypes
    idx = 0
    for label, coll in zip(labels, colls):
        for p in range(len(coeffs) * len(coll.groups)):
            pub = tables.pubs[idx]
            data = results[label][p].data
            for off, reg in ((pub["obs_offset"], data.observable_measurements), (pub["qpd_offset"], data.qpd_measurements)):
                want = np.ascontiguousarray(reg.array)
                got = np.frombuffer((ctypes.c_uint8 * want.size).from_address(sink.base + int(off)), dtype=np.uint8)
                assert np.array_equal(got, want.reshape(-1)), (label, p)
            idx += 1


def _with_duplicates(results, coefficients, moves):
    """``moves``: (label, source tuple, destination tuple) -- the destination's PUB objects are replaced by the source's,
    as when a caller runs each distinct subexperiment once (reference cutting_experiments.py:146-157 emits repeats)."""
    out = dict(results)
    for label, src, dst in moves:
        G = len(out[label]) // len(coefficients)
        pubs = list(out[label])
        pubs[dst * G: (dst + 1) * G] = pubs[src * G: (src + 1) * G]
        out[label] = PrimitiveResult(pubs)
    return out


def _reduce_on_host(tables, tally):
    """Product over subsystems / sum over tuples from a tally table, walking the tables like K2 does."""
    shots = tables.pubs["shots"].astype(np.float64)
    res = np.zeros(tables.n_obs)
    for i, c in enumerate(tables.coeffs):
        cur = np.ones(tables.n_obs)
        for lab in tables.labels:
            for k in range(tables.n_obs):
                e0, e1 = (int(tables.lookup_start[int(lab["lookup_base"]) + k + d]) for d in (0, 1))
                vals = []
                for slot in tables.lookup[e0:e1]:
                    p = int(lab["pub_base"]) + i * int(lab["num_groups"]) + int(slot["group"])
                    vals.append(tally[int(tables.pubs["tally_offset"][p]) + int(slot["term"])] / shots[p])
                cur[k] *= np.mean(vals)
        res += c * cur
    return res


@pytest.mark.parametrize("block", [0, 3, 7])
def test_duplicate_pubs_are_staged_and_tallied_once(lib, block):
    """PUBs that are the same buffers as an earlier PUB alias its tallies."""
    results, coefficients, observables = _cases.SMALL_CASES["molecular"]()
    dup = _with_duplicates(results, coefficients, [("S1", 0, 3), ("S2", 2, 4), ("S2", 2, 1)])
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)
    args = ([dup[label] for label in labels], colls, coeffs, 0, len(coeffs), len(observables[labels[0]]))
    tables, arrays = engine.build_host_tables(*args, block_pubs=block)
    plain, plain_arrays = engine.build_host_tables(*([results[label] for label in labels],) + args[1:], block_pubs=block)
    assert plain.k1_pubs is None and plain.dedup["dedup_ratio"] == 1.0
    G1, G2 = len(colls[1].groups), len(colls[2].groups)
    assert tables.k1_pubs is not None and len(tables.k1_pubs) == len(tables.pubs) - G1 - 2 * G2
    assert len(arrays) == 2 * len(tables.k1_pubs) and tables.data_bytes < plain.data_bytes and tables.n_tallies < plain.n_tallies
    assert tables.dedup["distinct_pubs"] == len(tables.k1_pubs) and tables.dedup["dedup_ratio"] > 1.0
    # a duplicate carries the offsets of its first occurrence
    b1 = int(tables.labels[1]["pub_base"])
    for f in ("obs_offset", "qpd_offset", "tally_offset", "shots", "group"):
        assert np.array_equal(tables.pubs[f][b1 + 3 * G1: b1 + 4 * G1], tables.pubs[f][b1: b1 + G1]), f
    # staged bytes + distinct PUBs -> tallies (C oracle) -> product/sum over ALL rows == the oracle on the duplicated input
    data = np.zeros(tables.data_bytes, dtype=np.uint8)
    engine.gather_arrays(tables, arrays, data.ctypes.data, 2)
    _, tally = c_oracle.pubs(data, tables.k1_pubs, tables.groups, tables.masks, tables.n_tallies, want_seq=False, n_threads=2)
    want = np.array(oracle.reconstruct_exact(dup, coefficients, observables))
    assert np.max(np.abs(_reduce_on_host(tables, tally) - want)) <= 1e-14 * max(1.0, float(np.abs(coeffs).sum()))


def test_same_address_but_different_shape_or_group_is_not_a_duplicate(lib):
    """Only identical (buffers, shape, group) triples are merged: a shared observable buffer with a different qpd
    buffer, or the same buffers under another group's masks, are tallied separately."""
    rng = np.random.default_rng(9)
    observables = {"A": PauliList(["ZI", "IZ", "XX"])}  # two groups: {ZI, IZ} and {XX}
    colls = [ObservableCollection(observables["A"])]
    assert len(colls[0].groups) == 2
    o = _cases.random_bits(rng, 50, 2)
    q1, q2 = _cases.random_bits(rng, 50, 3), _cases.random_bits(rng, 50, 3)
    pub = lambda obs, qpd: PubResult(DataBin(observable_measurements=BitArray(obs, 2), qpd_measurements=BitArray(qpd, 3)))  # noqa: E731
    shared = pub(o, q1)
    results = {"A": PrimitiveResult([shared, shared, pub(o, q2), pub(o, q1), shared, pub(o, q1)])}  # 3 tuples x 2 groups
    coefficients = [(0.5, WeightType.EXACT), (-0.25, WeightType.EXACT), (1.5, WeightType.EXACT)]
    tables, arrays = engine.build_host_tables([results["A"]], colls, np.array([0.5, -0.25, 1.5]), 0, 3, 3)
    # PUB 1 has PUB 0's buffers but group 1's masks; PUB 2 another qpd buffer; PUB 3 = buffers of PUB 1 (group 1) -> duplicate
    # of PUB 1; PUB 4 = PUB 0 (group 0) -> duplicate; PUB 5 = PUB 1/3 again
    assert tables.k1_pubs is not None and len(tables.k1_pubs) == 3
    data = np.zeros(tables.data_bytes, dtype=np.uint8)
    engine.gather_arrays(tables, arrays, data.ctypes.data, 1)
    _, tally = c_oracle.pubs(data, tables.k1_pubs, tables.groups, tables.masks, tables.n_tallies, want_seq=False)
    want = np.array(oracle.reconstruct_exact(results, coefficients, observables))
    assert np.max(np.abs(_reduce_on_host(tables, tally) - want)) <= 1e-14
