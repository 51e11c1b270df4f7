"""CPU checks of the bit-sliced kernel's logic: generation, NVRTC compilation for sm_100a, and lane-by-lane
emulation of the very same source against the oracle's exact tallies (bit-exact)."""

import numpy as np
import pytest

import _cases
import _golden
import _host_emu
from oracle import oracle
from qiskit_addon_cutting_b200 import _lib, engine
from qiskit_addon_cutting_b200.containers import BitArray, DataBin, PauliList, PrimitiveResult, PubResult, WeightType
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection


def _stage_host(results, coefficients, observables):
    if not isinstance(observables, dict):
        observables, results = {"A": observables}, {"A": results}
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)
    n_obs = len(observables[labels[0]])
    tables, arrays = engine.build_host_tables([results[label] for label in labels], colls, coeffs, 0, len(coeffs), n_obs)
    data = np.full(tables.data_bytes + 64, 0xA5, dtype=np.uint8)  # poison the padding: tails must be masked
    engine.gather_arrays(tables, arrays, data.ctypes.data, 1)
    want = np.concatenate(
        [t for label in labels for t in oracle.all_tallies(results, coefficients, observables)[label]]
    )
    return tables, data, want


@pytest.mark.parametrize("engine_kind", ["jit", "rt"])
@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_emulated_sliced_kernel_matches_oracle_on_golden_cases(lib, name, engine_kind):
    case = _golden.load_case(name)
    tables, data, want = _stage_host(case["results"], case["coefficients"], case["observables"])
    got, covered = _host_emu.emulate_tallies(tables, data, engine_kind)
    if not covered.any():
        pytest.skip("no group of this case is covered by the bit-sliced kernels")
    for p, pub in enumerate(tables.pubs):
        if covered[p]:
            sl = slice(int(pub["tally_offset"]), int(pub["tally_offset"]) + int(tables.groups[pub["group"]]["n_terms"]))
            assert np.array_equal(got[sl], want[sl]), (name, p)


@pytest.mark.parametrize("nbits,qbits", [(5, 3), (13, 1), (20, 1), (27, 2), (35, 8), (44, 3), (53, 1), (64, 5), (8, 12), (61, 20)])
@pytest.mark.parametrize("shots", [1, 31, 1024, 4097, 6200])
def test_emulated_sliced_kernel_widths_and_tails(lib, nbits, qbits, shots):
    """Every covered row width x qpd width, with tile/step boundaries: full steps, ragged tails, multi-tile."""
    rng = np.random.default_rng(nbits * 1000 + shots)
    # masks depend on the width only, so one emulator build serves every shot count
    observables = {"A": _cases.random_paulis(np.random.default_rng(nbits), 9, nbits, 1, nbits, "Z")}
    coefficients = [(1.0, WeightType.EXACT)]
    results = _cases.make_results(observables, 1, shots, qbits, seed=int(rng.integers(1 << 30)))
    tables, data, want = _stage_host(results, coefficients, observables)
    for engine_kind in ("jit", "rt"):
        got, covered = _host_emu.emulate_tallies(tables, data, engine_kind)
        assert covered.all(), "rows of 1..8 bytes must be covered by the bit-sliced kernels"
        assert np.array_equal(got, want), engine_kind


def test_nvrtc_compiles_golden_plans_for_sm100a(lib):
    compiled = 0
    for name in ("tutorial", "heavy_hex", "wide_dense", "molecular"):
        case = _golden.load_case(name)
        tables, _, _ = _stage_host(case["results"], case["coefficients"], case["observables"])
        for g in range(min(len(tables.groups), 2)):
            source = _lib.generate_source(tables.groups, tables.masks, g)
            if source:
                cubin = _lib.compile_source(source)
                assert cubin[:4] == b"\x7fELF" and len(cubin) > 1000
                compiled += 1
    assert compiled >= 4


def test_unsupported_widths_are_not_specialised(lib):
    groups = np.zeros(4, dtype=_lib.GROUP_DTYPE)
    groups[0] = (2, 33, 0, 0)    # 33-byte rows: generic kernel
    groups[1] = (2, 8, 66, 0)    # 8-byte rows
    groups[2] = (2, 16, 82, 0)   # 16-byte rows: wide-row specialised kernel
    groups[3] = (2, 32, 114, 0)  # 32-byte rows: wide-row specialised kernel, 2 warps per CTA
    masks = np.arange(178, dtype=np.uint8)
    assert _lib.generate_source(groups, masks, 0) == ""
    assert "struct Terms" in _lib.generate_source(groups, masks, 1)
    assert "qcut_lane_tile_wide" in _lib.generate_source(groups, masks, 2)
    assert "qcut_lane_tile_wide" in _lib.generate_source(groups, masks, 3)


def _guarded_copy(data: np.ndarray, used: int):
    """Copy `data[:used + DATA_SLACK]` so that it ENDS at an inaccessible guard page: any emulated load past the
    slack the C ABI promises (QCUT_DATA_SLACK) faults instead of silently reading a neighbour."""
    import ctypes
    import mmap

    page = mmap.PAGESIZE
    span = used + _lib.DATA_SLACK
    total = (span + page - 1) // page * page + page
    mm = mmap.mmap(-1, total, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, prot=mmap.PROT_READ | mmap.PROT_WRITE)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    start = total - page - span
    start -= start % 16  # keep the 16-byte alignment of the staged layout
    view = np.frombuffer(mm, dtype=np.uint8)
    view[start:start + used] = data[:used]
    view[start + used: total - page] = 0x5A
    libc = ctypes.CDLL(None)
    assert libc.mprotect(ctypes.c_void_p(base + total - page), ctypes.c_size_t(page), 0) == 0  # PROT_NONE
    return mm, view[start:]


@pytest.mark.parametrize("nbits,qbits,shots", [(20, 1, 5), (53, 3, 77), (35, 9, 1030), (64, 1, 4097), (13, 1, 2049), (5, 2, 4100)])
def test_emulated_loads_stay_inside_the_promised_slack(lib, nbits, qbits, shots):
    observables = {"A": _cases.random_paulis(np.random.default_rng(nbits), 9, nbits, 1, nbits, "Z")}
    coefficients = [(1.0, WeightType.EXACT)]
    results = _cases.make_results(observables, 1, shots, qbits, seed=3)
    tables, data, want = _stage_host(results, coefficients, observables)
    mm, guarded = _guarded_copy(data, tables.data_bytes)
    for engine_kind in ("jit", "rt"):
        got, covered = _host_emu.emulate_tallies(tables, guarded, engine_kind)
        assert covered.all() and np.array_equal(got, want)
    del guarded


@pytest.mark.parametrize("nbits,qbits", [(65, 1), (72, 3), (81, 1), (96, 2), (100, 9), (111, 1), (120, 4), (127, 8), (128, 1),
                                         (129, 1), (133, 2), (156, 1), (160, 3), (191, 1), (200, 9), (233, 1), (256, 1)])
@pytest.mark.parametrize("shots", [7, 1024, 1500, 4100])
def test_emulated_wide_row_kernel(lib, nbits, qbits, shots):
    """Rows of 9..32 bytes: shared-memory staged rows + parked planes (specialised kernels only)."""
    rng = np.random.default_rng(nbits * 31 + shots)
    observables = {"A": _cases.random_paulis(np.random.default_rng(nbits), 11, nbits, 1, min(nbits, 50), "Z")}
    coefficients = [(1.0, WeightType.EXACT)]
    results = _cases.make_results(observables, 1, shots, qbits, seed=int(rng.integers(1 << 30)))
    tables, data, want = _stage_host(results, coefficients, observables)
    mm, guarded = _guarded_copy(data, tables.data_bytes)
    got, covered = _host_emu.emulate_tallies(tables, guarded, "jit")
    assert covered.all(), "rows of 9..32 bytes must be covered by the specialised kernels"
    assert np.array_equal(got, want)
    del guarded


@pytest.mark.parametrize("shots", [1, 1024, 1500, 4097, 6200])
def test_emulated_split_engine_for_8_byte_rows(lib, monkeypatch, shots):
    """The experimental one-block-at-a-time engine (QCUT_JIT_SPLIT=1; off by default, measured slower): block-local terms,
    a few terms straddling the two halves of the row (their partial words are kept between the halves), identity."""
    monkeypatch.setenv("QCUT_JIT_SPLIT", "1")
    labels = []
    for q in range(64):  # single-qubit terms on every qubit: 64 measured bits -> 8-byte rows
        labels.append("".join("Z" if 63 - i == q else "I" for i in range(64)))
    for a, b in ((3, 4), (30, 31), (31, 32), (28, 40), (33, 60), (62, 63), (0, 63)):  # pairs, three of them cross
        labels.append("".join("Z" if 63 - i in (a, b) else "I" for i in range(64)))
    labels.append("I" * 64)
    observables = {"A": PauliList(labels)}
    results = _cases.make_results(observables, 1, shots, 3, seed=shots)
    tables, data, want = _stage_host(results, [(1.0, WeightType.EXACT)], observables)
    source = _lib.generate_source(tables.groups, tables.masks, 0)
    assert "SPLIT = 1" in source and "NKEEP = 3" in source and "qcut_lane_tile_split" in source
    got, covered = _host_emu.emulate_tallies(tables, data, "jit")
    assert covered.all() and np.array_equal(got, want)
    monkeypatch.delenv("QCUT_JIT_SPLIT")
    assert "qcut_lane_tile_split<Terms>" not in _lib.generate_source(tables.groups, tables.masks, 0).split("extern \"C\"", 1)[1]


def _heavy_hex_like_labels(n=63):
    """Single-Z on every qubit, ZZ on a chain plus a few longer bonds (two straddle the halves of the row), identity."""
    labels = ["".join("Z" if n - 1 - i == q else "I" for i in range(n)) for q in range(n)]
    bonds = [(q, q + 1) for q in range(0, n - 1, 1)] + [(2, 17), (9, 40), (20, 35), (45, 60), (31, 32)]
    labels += ["".join("Z" if n - 1 - i in ab else "I" for i in range(n)) for ab in bonds]
    labels.append("I" * n)
    return labels


@pytest.mark.parametrize("n_warps", [1, 3])
def test_emulated_pipelined_stream_for_8_byte_rows(lib, monkeypatch, n_warps):
    """QCUT_JIT_PIPE=1 (experimental, off by default).  The register-level software pipeline (qcut_warp_pipe): a warp walks its tiles first, first + stride, ... as one
    flat stream of steps and the rows of the next step -- or of the first step of its next tile -- are loaded from
    inside the term network.  PUBs with full tiles, ragged tails, a single shot, no shots; several warps."""
    monkeypatch.setenv("QCUT_JIT_PIPE", "1")
    observables = {"A": PauliList(_heavy_hex_like_labels())}
    shots_per_pub = [8192, 5000, 1, 4096, 1024, 0, 3071, 12289, 2048]
    coefficients = [(1.0, WeightType.EXACT)] * len(shots_per_pub)
    rng = np.random.default_rng(7)
    pubs = []
    for i, shots in enumerate(shots_per_pub):
        one = _cases.make_results(observables, 1, shots, 1 + (i % 3 == 2) * 8, seed=100 + i)
        pubs.append(one["A"][0])
    results = {"A": PrimitiveResult(pubs)}
    tables, data, want = _stage_host(results, coefficients, observables)
    source = _lib.generate_source(tables.groups, tables.masks, 0)
    assert "PIPE = 1" in source and "qcut_warp_pipe<Terms>" in source.split("extern \"C\"", 1)[1]
    emu = _host_emu.compile_emulator(source)
    tiles = np.array([(p, t) for p, pub in enumerate(tables.pubs) for t in range(-(-int(pub["shots"]) // _lib.TILE_SHOTS))],
                     dtype=np.uint32)
    mm, guarded = _guarded_copy(data, tables.data_bytes)
    got = np.zeros(tables.n_tallies, dtype=np.uint64)
    pub_table = np.ascontiguousarray(tables.pubs)
    for warp in range(n_warps):
        for lane in range(32):
            assert emu.qcut_emu_lane_stream(guarded.ctypes.data, pub_table.ctypes.data, tiles.ctypes.data, len(tiles), warp,
                                            n_warps, lane, got.ctypes.data) == 1
    assert np.array_equal(got.view(np.int64), want)
    del guarded
    # dense masks: the pipelined form is not offered (every term would be built from scratch)
    dense = {"A": _cases.random_paulis(np.random.default_rng(1), 40, 64, 20, 40, "Z")}
    t2, _, _ = _stage_host(_cases.make_results(dense, 1, 10, 1, seed=1), [(1.0, WeightType.EXACT)], dense)
    assert "PIPE = 0" in _lib.generate_source(t2.groups, t2.masks, 0)


@pytest.mark.parametrize("nbits,shots", [(63, 8192), (63, 1500), (30, 4097), (12, 6200), (7, 9000), (44, 1)])
def test_emulated_shared_memory_counters(lib, monkeypatch, nbits, shots):
    """QCUT_JIT_SACC=1: the packed counters of a tile live in the thread's column of a shared-memory array, sixteen terms
    per 128-bit load / store (sparse masks, rows of up to 8 bytes)."""
    monkeypatch.setenv("QCUT_JIT_SACC", "1")
    labels = _heavy_hex_like_labels(63)
    observables = {"A": PauliList(sorted({lab[63 - nbits:] for lab in labels}))}
    results = _cases.make_results(observables, 1, shots, 2, seed=shots + nbits)
    tables, data, want = _stage_host(results, [(1.0, WeightType.EXACT)], observables)
    source = _lib.generate_source(tables.groups, tables.masks, 0)
    assert "SACC = 1" in source and "qcut_lane_tile_sacc<Terms::NB, Terms>" in source.split("extern \"C\"", 1)[1]
    got, covered = _host_emu.emulate_tallies(tables, data, "jit")
    assert covered.all() and np.array_equal(got, want)


@pytest.mark.parametrize("shots", [1, 1024, 1500, 4097, 8192])
def test_emulated_interleaved_engine_for_8_byte_rows(lib, monkeypatch, shots):
    """QCUT_JIT_IL=1: the transposition of block 1 is spread, one pair of rows at a time, between the terms of block 0
    (the five stages commute, qcut_tp); full steps take that path, tails and wide qpd rows the all-at-once one."""
    monkeypatch.setenv("QCUT_JIT_IL", "1")
    observables = {"A": PauliList(_heavy_hex_like_labels())}
    for qbits in (3, 11):
        results = _cases.make_results(observables, 1, shots, qbits, seed=shots + qbits)
        tables, data, want = _stage_host(results, [(1.0, WeightType.EXACT)], observables)
        source = _lib.generate_source(tables.groups, tables.masks, 0)
        assert "IL = 1" in source and "qcut_lane_tile_il<Terms>" in source.split("extern \"C\"", 1)[1]
        got, covered = _host_emu.emulate_tallies(tables, data, "jit")
        assert covered.all() and np.array_equal(got, want)


@pytest.mark.parametrize("nbits", [127, 128, 190, 256])
@pytest.mark.parametrize("shots", [1, 1024, 1500, 4100])
def test_emulated_slab_engine_for_16_24_32_byte_rows(lib, nbits, shots):
    """Rows of 16 / 24 / 32 bytes with a local observable (what an un-partitioned 127-qubit register looks like): the row is
    processed in 8-byte slabs on the narrow-row engine (qcut_lane_tile_slab) -- slab-local terms, terms that straddle two
    or three slabs (partial words kept between the passes), identity, ragged tails, 1- and 2-byte qpd rows."""
    labels = ["".join("Z" if nbits - 1 - i == q else "I" for i in range(nbits)) for q in range(nbits)]
    bonds = [(q, q + 1) for q in range(0, nbits - 1, 2 if nbits <= 128 else 3)] + [(60, 70), (63, 64), (5, 126), (100, nbits - 1)]
    if nbits > 128:
        bonds += [(120, 135), (10, 70, 150)]
    labels += ["".join("Z" if nbits - 1 - i in qs else "I" for i in range(nbits)) for qs in bonds]
    labels.append("I" * nbits)
    observables = {"A": PauliList(labels)}
    for qbits in (3, 11):
        results = _cases.make_results(observables, 1, shots, qbits, seed=shots + qbits + nbits)
        tables, data, want = _stage_host(results, [(1.0, WeightType.EXACT)], observables)
        source = _lib.generate_source(tables.groups, tables.masks, 0)
        assert "SLAB = 1" in source and "qcut_lane_tile_slab<Terms::NB, Terms>" in source.split("extern \"C\"", 1)[1]
        mm, guarded = _guarded_copy(data, tables.data_bytes)
        got, covered = _host_emu.emulate_tallies(tables, guarded, "jit")
        assert covered.all() and np.array_equal(got, want)
        del guarded


def test_emulated_switch_kernels_of_a_molecular_plan(lib):
    """Small groups share one kernel per row width, their networks behind a switch on the group's variant index
    (generate_switch_source).  The same source, compiled for the host, is run lane by lane for every PUB of a
    molecular-style case (many groups of a few terms on 1- and 2-byte rows, ragged shot counts) and must reproduce the
    oracle's tallies; the packed-counter layout differs per variant (ceil(T / 4) registers)."""
    rng = np.random.default_rng(23)

    def strings(n_qubits, count):  # weight-1..3 XYZ strings: qubit-wise commuting groups of a few terms each
        labels = set()
        while len(labels) < count:
            chars = ["I"] * n_qubits
            for q in rng.choice(n_qubits, size=int(rng.integers(1, 4)), replace=False):
                chars[q] = "XYZ"[int(rng.integers(3))]
            labels.add("".join(chars))
        return sorted(labels)

    def families(n_qubits, n_families, per_family):  # each family: sub-strings of one long XYZ string -> one wide group
        labels = []
        for _ in range(n_families):
            base = ["XYZ"[int(rng.integers(3))] for _ in range(n_qubits)]
            for _ in range(per_family):
                keep = rng.random(n_qubits) < 0.7
                labels.append("".join(c if k else "I" for c, k in zip(base, keep)))
        return labels

    # three subsystems: 7 qubits (1-byte rows), 12 qubits (2-byte rows), 60 qubits (8-byte rows), 50 observables each
    observables = {"A": PauliList(strings(7, 50)), "B": PauliList(strings(12, 50)), "C": PauliList(families(60, 5, 10))}
    results = _cases.make_results(observables, 2, 4500, 2, seed=5)
    coefficients = [(0.5, WeightType.EXACT), (-0.25, WeightType.EXACT)]
    tables, data, want = _stage_host(results, coefficients, observables)
    widths, counts = np.unique(tables.groups["nb_obs"], return_counts=True)
    assert widths.tolist() == [1, 2, 8] and (counts >= 4).all() and int(tables.groups["n_terms"].max()) <= 32
    sources, k = {}, 0
    while True:
        src, kernel_of, variant_of = _lib.plan_kernel_source(tables.groups, tables.masks, k)
        if not src:
            break
        sources[k] = src
        k += 1
    assert sources and all("struct Net" in s and "switch (variant)" in s for s in sources.values()), "switch kernels expected"
    assert (kernel_of >= 0).all()
    emus = {kk: _host_emu.compile_emulator(src) for kk, src in sources.items()}
    got = np.zeros(tables.n_tallies, dtype=np.int64)
    odd = np.zeros(64, dtype=np.uint32)
    for pub in tables.pubs:
        g = int(pub["group"])
        grp = tables.groups[g]
        nb, nbq, shots, T = int(grp["nb_obs"]), int(pub["nb_qpd"]), int(pub["shots"]), int(grp["n_terms"])
        for shot0 in range(0, shots, _lib.TILE_SHOTS):
            ns = min(_lib.TILE_SHOTS, shots - shot0)
            total = np.zeros(T, dtype=np.int64)
            for lane in range(32):
                n = emus[int(kernel_of[g])].qcut_emu_lane_tile_variant(
                    int(variant_of[g]), data.ctypes.data + int(pub["obs_offset"]) + shot0 * nb,
                    data.ctypes.data + int(pub["qpd_offset"]) + shot0 * nbq, nbq, ns, lane, odd.ctypes.data)
                assert n == T
                total += odd[:T]
            got[int(pub["tally_offset"]): int(pub["tally_offset"]) + T] += ns - 2 * total
    assert np.array_equal(got, want)
