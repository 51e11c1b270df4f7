"""Parity tests proper (B200): the CUDA path, through the public API and the C ABI, against the oracle.

Integer tallies must be bit-exact.  Expectation values are compared three ways:
  * vs the oracle with E = tally/S ("exact"): the only differences allowed are fp64 summation-order
    effects of K2, bounded here by 4 ulp of the largest partial sum -- in practice 0 ulp because K2
    keeps the reference's order for C <= 1024;
  * vs what the reference's own source produced (golden fixtures): within the reference's
    self-noise from accumulating +-fl(1/S) shot by shot (1e-13 * kappa for S <= 4100);
  * bit-identical to the reference when S is a power of two.
"""

import numpy as np
import pytest
import torch

import _cases
import _golden
from oracle import oracle
from qiskit_addon_cutting_b200 import (
    BitArray, DataBin, PauliList, PrimitiveResult, PubResult, QuasiDistribution, SamplerResult, WeightType,
    _lib, engine, reconstruct_expectation_values,
)
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection

pytestmark = pytest.mark.gpu

KAT = _golden.load_kat()


def _hex(values):
    return [float(v).hex() for v in values]


def _device_tallies(results, coefficients, observables, flags=0):
    if not isinstance(observables, dict):
        observables, results = {"A": observables}, {"A": results}
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)
    batch = engine.stage_v2([results[label] for label in labels], colls, coeffs, 0, len(coeffs),
                            len(observables[labels[0]]), flags=flags)
    batch.run()
    torch.cuda.synchronize()
    want = np.concatenate([t for label in labels for t in oracle.all_tallies(results, coefficients, observables)[label]])
    return batch, batch.tallies.cpu().numpy()[: batch.tables.n_tallies], want


def test_library_loaded_and_device_visible(lib):
    assert lib.qcut_device_count() >= 1
    assert torch.cuda.get_device_capability(0)[0] >= 10, "these kernels are built for sm_100a"


def test_reference_v2_golden_vector(lib):
    obs = BitArray(np.array([0, 0, 0, 1, 0, 0, 2, 3, 1, 0], dtype=np.uint8).reshape(-1, 1), 2)
    qpd = BitArray(np.array([0, 1, 1, 3, 2, 0, 1, 3, 0, 2], dtype=np.uint8).reshape(-1, 1), 2)
    results = PrimitiveResult([PubResult(DataBin(observable_measurements=obs, qpd_measurements=qpd))])
    out = reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], PauliList(["II", "IZ", "ZI", "ZZ"]))
    assert isinstance(out, list) and all(isinstance(v, np.float64) for v in out)
    assert out == pytest.approx([0.0, -0.6, 0.0, -0.2])


def test_reference_v1_trivial_and_truth_table(lib):
    results = SamplerResult(quasi_dists=[QuasiDistribution({"0": 1.0})], metadata=[{}])
    assert reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], PauliList(["ZZ"])) == [1.0]
    # _process_outcome truth table (cog XZ: IZ, XI, XZ) through the V1 kernel: one outcome of weight 1
    for outcome, expected in KAT["process_outcome"]["table"].items():
        for key in (outcome, f"0b{outcome}", int(outcome, 2), hex(int(outcome, 2))):
            results = SamplerResult(quasi_dists=[{key: 1.0}])
            got = reconstruct_expectation_values(results, [(1.0, WeightType.EXACT)], PauliList(["IZ", "XI", "XZ"]))
            assert got == expected, (key, got)


def test_v1_case_bit_identical_to_reference(lib):
    case = _golden.load_v1_case()
    out = reconstruct_expectation_values(case["results"], case["coefficients"], case["observables"])
    assert _hex(out) == _hex(case["reference_expvals"])


@pytest.mark.parametrize("flags", [0, _lib.PLAN_NO_JIT, _lib.PLAN_GENERIC_ONLY, _lib.PLAN_STAGED],
                         ids=["jit", "sliced_rt", "generic", "jit_staged"])
@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_tallies_bit_exact_on_golden_cases(lib, name, flags):
    case = _golden.load_case(name)
    batch, got, want = _device_tallies(case["results"], case["coefficients"], case["observables"], flags)
    assert np.array_equal(got, want)
    kernels = set(batch.plan.kernels())
    if flags == _lib.PLAN_GENERIC_ONLY:
        assert kernels == {0}
    if flags == _lib.PLAN_NO_JIT:
        assert max(kernels) <= 1


@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_expvals_on_golden_cases(lib, name):
    case = _golden.load_case(name)
    out = np.array(reconstruct_expectation_values(case["results"], case["coefficients"], case["observables"]))
    kappa = max(1.0, sum(abs(c) for c, _ in case["coefficients"]))
    exact = np.array(oracle.reconstruct_exact(case["results"], case["coefficients"], case["observables"]))
    assert np.max(np.abs(out - exact)) <= 4 * np.finfo(np.float64).eps * kappa
    assert _hex(out) == _hex(exact), "K2 keeps the reference's summation order for C <= 1024"
    assert np.max(np.abs(out - np.array(case["reference_expvals"]))) <= 1e-13 * kappa
    if name == "pow2":
        assert _hex(out) == _hex(case["reference_expvals"])


@pytest.mark.parametrize("nbits,qbits", [(3, 1), (8, 8), (9, 2), (16, 1), (17, 3), (24, 9), (31, 1), (33, 4),
                                          (47, 1), (56, 8), (64, 1), (65, 2), (100, 1), (127, 8), (133, 2), (156, 1),
                                          (200, 17), (256, 3), (264, 1)])
@pytest.mark.parametrize("shots", [1, 100, 4096, 5000])
def test_tallies_every_row_width(lib, nbits, qbits, shots):
    """Row widths 1..33 bytes (bit-sliced kernels up to 32 bytes, generic beyond), qpd widths 1..3 bytes, ragged
    shot counts."""
    rng = np.random.default_rng(nbits * 7919 + shots)
    observables = {"A": _cases.random_paulis(rng, 20, nbits, 1, min(nbits, 40), "Z")}
    results = _cases.make_results(observables, 2, shots, qbits, seed=int(rng.integers(1 << 30)))
    _, got, want = _device_tallies(results, [(0.5, WeightType.EXACT), (-0.5, WeightType.EXACT)], observables)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("staged", [False, True])
@pytest.mark.parametrize("nbits,n_terms", [(5, 16), (12, 16), (23, 40), (32, 16), (39, 16), (48, 40), (50, 16), (64, 40), (64, 16)])
def test_staged_stream_of_steps_matches_direct_loads(lib, monkeypatch, nbits, n_terms, staged):
    """The bulk-copy staged stream (one step ahead through shared memory, chosen automatically for narrow rows with few
    terms, QCUT_PLAN_STAGED elsewhere) against the direct-load engine: every layout (1..8-byte rows), PUBs of many
    tiles, ragged tails, a single shot and empty PUBs in between, wide qpd rows."""
    monkeypatch.setenv("QCUT_JIT_AUTO_STAGED", "1" if staged else "0")
    rng = np.random.default_rng(nbits)
    # (up to 32 terms: kernel shared by the small groups of one row width, networks behind a switch; more: own kernel)
    observables = {"A": _cases.random_paulis(rng, n_terms, nbits, 1, min(nbits, 6), "Z")}
    pubs = []
    for i, shots in enumerate([9000, 1, 4096, 0, 12289, 1024, 5000, 3071]):
        one = _cases.make_results(observables, 1, shots, 1 + 8 * (i % 4 == 3), seed=nbits * 100 + i)
        pubs.append(one["A"][0])
    results = {"A": PrimitiveResult(pubs)}
    coefficients = [(1.0, WeightType.EXACT)] * len(pubs)
    batch, got, want = _device_tallies(results, coefficients, observables, _lib.PLAN_STAGED if staged else 0)
    assert np.array_equal(got, want)
    source = batch.plan.source(_lib.KERNEL_JIT0)
    if "qcut_variant_0" not in source:  # (fewer than 12 distinct terms share a multi-group kernel: direct loads)
        assert ("qcut_warp_stream" in source.split('extern "C"', 1)[1]) == staged


@pytest.mark.parametrize("nbits", [127, 190, 256])
def test_slab_engine_for_16_24_32_byte_rows(lib, monkeypatch, nbits):
    """Local observable on a 127- / 190- / 256-bit register (16 / 24 / 32-byte rows): processed in 8-byte slabs on the
    narrow-row engine, bit-exact against the oracle and against the shared-memory wide-row engine (QCUT_JIT_SLAB=0)."""
    step = 2 if nbits <= 128 else 3
    labels = ["".join("Z" if nbits - 1 - i == q else "I" for i in range(nbits)) for q in range(nbits)]
    bonds = [(q, q + 1) for q in range(0, nbits - 1, step)] + [(60, 70), (63, 64), (5, 126), (100, nbits - 1)]
    if nbits > 128:
        bonds += [(120, 135), (10, 70, 150)]
    labels += ["".join("Z" if nbits - 1 - i in qs else "I" for i in range(nbits)) for qs in bonds]
    labels.append("I" * nbits)
    observables = {"A": PauliList(labels)}
    pubs = []
    for i, shots in enumerate([9000, 1, 4096, 0, 12289, 1024, 5000, 3071]):
        one = _cases.make_results(observables, 1, shots, 1 + 8 * (i % 4 == 3), seed=nbits * 100 + i)
        pubs.append(one["A"][0])
    results = {"A": PrimitiveResult(pubs)}
    coefficients = [(1.0, WeightType.EXACT)] * len(pubs)
    batch, got, want = _device_tallies(results, coefficients, observables)
    assert np.array_equal(got, want)
    assert "qcut_lane_tile_slab" in batch.plan.source(_lib.KERNEL_JIT0).split('extern "C"', 1)[1]
    monkeypatch.setenv("QCUT_JIT_SLAB", "0")
    batch2, got2, _ = _device_tallies(results, coefficients, observables, _lib.PLAN_STAGED)  # (other flags: not the cached plan)
    assert np.array_equal(got2, want)
    assert "qcut_lane_tile_wide" in batch2.plan.source(_lib.KERNEL_JIT0).split('extern "C"', 1)[1]


def test_rows_of_17_to_32_bytes_use_the_specialised_kernel(lib):
    """A 156-qubit register (20-byte rows) is handled by the wide-row engine, 33-byte rows by the generic kernel."""
    for nbits, specialised in ((156, True), (256, True), (264, False)):
        rng = np.random.default_rng(nbits)
        observables = {"A": _cases.random_paulis(rng, 30, nbits, nbits // 2, nbits, "Z")}  # every qubit is measured
        results = _cases.make_results(observables, 1, 6000, 1, seed=nbits)
        batch, got, want = _device_tallies(results, [(1.0, WeightType.EXACT)], observables)
        assert np.array_equal(got, want)
        assert batch.tables.groups["nb_obs"].tolist() == [(nbits + 7) // 8]
        assert (max(batch.plan.kernels()) >= _lib.KERNEL_JIT0) == specialised
        if specialised:
            assert "qcut_lane_tile_wide" in batch.plan.source(_lib.KERNEL_JIT0)


def test_empty_pubs_and_zero_shots(lib):
    """A PUB with zero shots contributes E = 0 (the reference's shot loop does not execute)."""
    observables = {"A": PauliList(["ZI", "IZ"]), "B": PauliList(["II", "ZZ"])}
    rng = np.random.default_rng(5)

    def pub(shots, bits):
        return PubResult(DataBin(observable_measurements=BitArray(_cases.random_bits(rng, shots, bits), bits),
                                 qpd_measurements=BitArray(_cases.random_bits(rng, shots, 2), 2)))

    results = {"A": PrimitiveResult([pub(0, 2), pub(7, 2)]), "B": PrimitiveResult([pub(5, 2), pub(0, 2)])}
    coefficients = [(0.75, WeightType.EXACT), (-0.25, WeightType.EXACT)]
    out = reconstruct_expectation_values(results, coefficients, observables)
    assert _hex(out) == _hex(oracle.reconstruct_exact(results, coefficients, observables))
    assert _hex(out) == _hex(oracle.reconstruct_literal(results, coefficients, observables)) or np.allclose(
        out, oracle.reconstruct_literal(results, coefficients, observables), atol=1e-15)


def test_many_terms_group_uses_specialised_kernel_and_matches(lib):
    """cfg4 / cfg3 shapes at small size: > 12 terms on 8-byte rows -> NVRTC-specialised kernel."""
    for name in ("heavy_hex", "wide_dense"):
        results, coefficients, observables = _cases.SMALL_CASES[name]()
        batch, got, want = _device_tallies(results, coefficients, observables)
        assert max(batch.plan.kernels()) >= _lib.KERNEL_JIT0
        assert "struct Terms" in batch.plan.source(_lib.KERNEL_JIT0)
        assert np.array_equal(got, want)


def test_large_chunked_reduction_matches_oracle(lib):
    """C > 1024 coefficient tuples: K2 runs several chunks; result within summation-order noise of the oracle."""
    rng = np.random.default_rng(77)
    observables = {"A": _cases.random_paulis(rng, 5, 8, 1, 4, "Z"), "B": _cases.random_paulis(rng, 5, 6, 1, 4, "Z")}
    C = 2500
    coefficients = _cases.random_coefficients(rng, C, kappa=81.0)
    results = _cases.make_results(observables, C, 64, 2, seed=78)
    out = np.array(reconstruct_expectation_values(results, coefficients, observables))
    want = np.array(oracle.reconstruct_exact(results, coefficients, observables))
    assert np.max(np.abs(out - want)) <= 1e-12


@pytest.mark.parametrize("n_labels,C", [(1, 3), (3, 40), (5, 7), (6, 1500)])
def test_reduction_for_any_number_of_subsystems(lib, n_labels, C):
    """K2 runs straight-line code for 1..4 subsystems and the general form beyond; both chunkings (C <= / > 1024)."""
    rng = np.random.default_rng(100 * n_labels + C)
    observables = {f"L{j}": _cases.random_paulis(rng, 6, 5 + j, 0, 4) for j in range(n_labels)}
    coefficients = _cases.random_coefficients(rng, C, kappa=27.0)
    results = _cases.make_results(observables, C, 48, 2, seed=int(rng.integers(1 << 30)))
    out = np.array(reconstruct_expectation_values(results, coefficients, observables))
    want = np.array(oracle.reconstruct_exact(results, coefficients, observables))
    if C <= 1024:
        assert _hex(out) == _hex(want)
    else:
        assert np.max(np.abs(out - want)) <= 1e-13 * 27.0


@pytest.mark.parametrize("name", ["molecular", "odd_widths", "wire"])
def test_block_staging_many_blocks(lib, monkeypatch, name):
    """Tiny staging blocks (every label cut into several device allocations): same tallies, same result."""
    results, coefficients, observables = _cases.SMALL_CASES[name]()
    monkeypatch.setattr(engine, "_BLOCK_PUBS", 2)
    monkeypatch.setattr(engine, "_SIMPLE_PUBS", 0)  # small inputs normally take the one-copy path: force the pipeline
    batch, got, want = _device_tallies(results, coefficients, observables)
    assert engine.last_stage_trace["path"] == "pipelined"
    assert len(batch.data.blocks) > len(observables if isinstance(observables, dict) else [0])
    assert np.array_equal(got, want)
    out = reconstruct_expectation_values(results, coefficients, observables)
    monkeypatch.undo()
    assert _hex(out) == _hex(reconstruct_expectation_values(results, coefficients, observables))


def test_concurrent_calls_from_python_threads(lib):
    """The reference is pure; ours must at least be safe to call from several threads (shared pinned ring, plan
    cache and side streams are serialised internally): every thread gets the result of its own inputs."""
    import threading

    cases = [_cases.SMALL_CASES[name]() for name in ("molecular", "odd_widths", "wire", "heavy_hex")]
    want = [_hex(reconstruct_expectation_values(*c)) for c in cases]
    got: list = [None] * (3 * len(cases))
    errors: list = []

    def work(slot):
        try:
            for _ in range(3):
                got[slot] = _hex(reconstruct_expectation_values(*cases[slot % len(cases)]))
        except BaseException as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(got))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert got == [want[i % len(cases)] for i in range(len(got))]


def _with_duplicates(results, coefficients, moves):
    out = dict(results)
    for label, src, dst in moves:
        G = len(out[label]) // len(coefficients)
        pubs = list(out[label])
        pubs[dst * G: (dst + 1) * G] = pubs[src * G: (src + 1) * G]
        out[label] = PrimitiveResult(pubs)
    return out


@pytest.mark.parametrize("pipelined", [False, True])
def test_duplicate_pubs_are_staged_and_tallied_once(lib, monkeypatch, pipelined):
    """SURVEY section 8f row 4 (reference cutting_experiments.py:146-157 emits repeated subexperiments): PUBs that are
    the same buffers as an earlier PUB are detected at ingest; fewer bytes cross PCIe, tallies of the distinct PUBs are
    bit-exact and the expectation values are bit-identical to the exact oracle on the duplicated input."""
    results, coefficients, observables = _cases.SMALL_CASES["molecular"]()
    dup = _with_duplicates(results, coefficients, [("S1", 0, 3), ("S2", 2, 4), ("S2", 2, 1), ("S0", 1, 2)])
    if pipelined:
        monkeypatch.setattr(engine, "_BLOCK_PUBS", 3)
        monkeypatch.setattr(engine, "_SIMPLE_PUBS", 0)
    labels = list(observables)
    colls = [ObservableCollection(observables[label]) for label in labels]
    coeffs = np.array([c for c, _ in coefficients], dtype=np.float64)
    plain = engine.stage_v2([results[k] for k in labels], colls, coeffs, 0, len(coeffs), len(observables[labels[0]]))
    batch = engine.stage_v2([dup[k] for k in labels], colls, coeffs, 0, len(coeffs), len(observables[labels[0]]))
    assert engine.last_stage_trace["path"] == ("pipelined" if pipelined else "simple")
    t = batch.tables
    assert t.k1_pubs is not None and len(t.k1_pubs) < len(t.pubs) and t.data_bytes < plain.tables.data_bytes
    assert t.dedup["dedup_ratio"] > 1.15 and plain.tables.dedup["dedup_ratio"] == 1.0
    batch.run()
    torch.cuda.synchronize()
    # tallies of the distinct PUBs, in staging order, against the oracle
    every = [x for k in labels for x in oracle.all_tallies(dup, coefficients, observables)[k]]
    first = sorted({int(o) for o in t.pubs["tally_offset"]})
    want = np.concatenate([every[int(np.flatnonzero(t.pubs["tally_offset"] == o)[0])] for o in first])
    assert np.array_equal(batch.tallies.cpu().numpy()[: t.n_tallies], want)
    got = reconstruct_expectation_values(dup, coefficients, observables)
    assert _hex(got) == _hex(oracle.reconstruct_exact(dup, coefficients, observables))
    monkeypatch.setattr(engine, "_DEDUP", False)
    assert _hex(got) == _hex(reconstruct_expectation_values(dup, coefficients, observables))


# ---- rounding="reference": the reference's own fp64 arithmetic, bit for bit -------------------------------------
@pytest.mark.parametrize("name", _golden.CASE_NAMES)
def test_reference_rounding_is_bit_identical_to_the_reference(lib, name):
    """Fixtures hold what the reference's own source text returned (tests/golden/make_golden.py); shot counts are not
    powers of two, so the reference's shot-by-shot accumulation differs from tally/S in the last bits -- the
    reference-rounding mode reproduces exactly those bits (cutting_reconstruction.py:159,168)."""
    case = _golden.load_case(name)
    out = reconstruct_expectation_values(case["results"], case["coefficients"], case["observables"], rounding="reference")
    assert _hex(out) == _hex(case["reference_expvals"])


def test_reference_rounding_for_every_row_width_and_many_tuples(lib, monkeypatch):
    """Widths 1..33 bytes (all word-count variants of the chain kernel), multi-slot means, zero-shot PUBs, C > 1024
    (sequential K2), and the environment-variable default: bit-identical to the oracle's literal restatement, which
    tests/test_oracle.py pins to the reference's text."""
    rng = np.random.default_rng(4242)
    for nbits, qbits, shots in ((3, 1, 100), (20, 11, 77), (33, 2, 130), (64, 8, 257), (100, 1, 90), (264, 3, 50)):
        observables = {"A": _cases.random_paulis(rng, 9, nbits, 1, min(nbits, 30), "Z"),
                       "B": _cases.random_paulis(rng, 9, 5, 0, 5)}
        coefficients = _cases.random_coefficients(rng, 5)
        results = _cases.make_results(observables, 5, lambda label, i, g: shots + i if i != 2 else 0, qbits,
                                      seed=int(rng.integers(1 << 30)))
        out = reconstruct_expectation_values(results, coefficients, observables, rounding="reference")
        assert _hex(out) == _hex(oracle.reconstruct_literal(results, coefficients, observables)), nbits
    # duplicates in the observable list -> multi-slot lookups (np.mean over slots), identity-only groups
    results, coefficients, observables = _cases.SMALL_CASES["odd_widths"]()
    assert _hex(reconstruct_expectation_values(results, coefficients, observables, rounding="reference")) == \
        _hex(oracle.reconstruct_literal(results, coefficients, observables))
    # more than 1024 tuples: the default mode sums chunks, the reference mode stays strictly sequential
    observables = {"A": _cases.random_paulis(rng, 5, 8, 1, 4, "Z"), "B": _cases.random_paulis(rng, 5, 6, 1, 4, "Z")}
    C = 2300
    coefficients = _cases.random_coefficients(rng, C, kappa=81.0)
    results = _cases.make_results(observables, C, 21, 2, seed=79)
    want = oracle.reconstruct_literal(results, coefficients, observables)
    monkeypatch.setenv("QCUT_ROUNDING", "reference")
    assert _hex(reconstruct_expectation_values(results, coefficients, observables)) == _hex(want)
    monkeypatch.delenv("QCUT_ROUNDING")
    assert np.max(np.abs(np.array(reconstruct_expectation_values(results, coefficients, observables)) - want)) <= 1e-12
    with pytest.raises(ValueError):
        reconstruct_expectation_values(results, coefficients, observables, rounding="fast")


def test_background_specialisation_has_no_cold_start(lib, monkeypatch, tmp_path):
    """Public calls create their plan with QCUT_PLAN_ASYNC_JIT: NVRTC runs on a background thread and the call is
    served by the ahead-of-time kernels meanwhile -- same tallies, same expectation values, no wait for the compiler."""
    import time

    monkeypatch.setenv("QCUT_CACHE_DIR", str(tmp_path))  # empty kernel cache: a real cold start
    monkeypatch.setattr(engine, "_SMALL_CALL_BYTES", 0)   # (small calls would not use NVRTC at all)
    rng = np.random.default_rng(20251)
    observables = {"A": _cases.random_paulis(rng, 40, 64, 1, 3, "Z"), "B": _cases.random_paulis(rng, 40, 60, 1, 3, "Z")}
    coefficients = _cases.random_coefficients(rng, 6)
    results = _cases.make_results(observables, 6, 300, 2, seed=5)
    want = _hex(oracle.reconstruct_exact(results, coefficients, observables))
    labels = list(observables)
    colls = [ObservableCollection(observables[k]) for k in labels]
    coeffs = np.array([c for c, _ in coefficients])
    t0 = time.perf_counter()
    batch = engine.stage_v2([results[k] for k in labels], colls, coeffs, 0, 6, 40, public_call=True)
    first = _hex(batch.run().cpu().numpy())
    first_s = time.perf_counter() - t0
    state_at_first_call = batch.plan.jit_state()
    kernels_first = set(batch.plan.kernels()) if state_at_first_call == 1 else None
    assert first == want
    batch.plan.wait()
    waited_s = time.perf_counter() - t0
    assert batch.plan.jit_state() == 2 and max(batch.plan.kernels()) >= _lib.KERNEL_JIT0
    assert _hex(reconstruct_expectation_values(results, coefficients, observables)) == want  # now on the specialised kernels
    if kernels_first is not None:  # the compiler was still running when the first call returned
        assert max(kernels_first) <= _lib.KERNEL_SLICED_RT
    print(f"\\nfirst call {1e3 * first_s:.1f} ms (jit state {state_at_first_call}), specialised kernels ready after {waited_s:.2f} s")
    # a plan destroyed while its compilation is in flight stops the compiler and does not hang
    groups = batch.tables.groups.copy()
    masks = batch.tables.masks.copy()
    masks[0] ^= 1  # another mask set -> another compilation
    monkeypatch.setenv("QCUT_CACHE_DIR", str(tmp_path / "other"))
    t0 = time.perf_counter()
    plan = engine.Plan(groups, masks, flags=_lib.PLAN_ASYNC_JIT)
    created_s = time.perf_counter() - t0
    del plan
    assert created_s < 0.5, "plan creation must not wait for NVRTC"


def test_small_public_calls_never_invoke_nvrtc(lib):
    """Launch-bound calls (less than 64 MiB staged) run on the ahead-of-time kernels: nothing to compile."""
    results, coefficients, observables = _cases.SMALL_CASES["heavy_hex"]()
    labels = list(observables)
    colls = [ObservableCollection(observables[k]) for k in labels]
    coeffs = np.array([c for c, _ in coefficients])
    batch = engine.stage_v2([results[k] for k in labels], colls, coeffs, 0, len(coeffs), len(observables["A"]), public_call=True)
    assert batch.plan.jit_state() == 0 and max(batch.plan.kernels()) <= _lib.KERNEL_SLICED_RT
    assert _hex(batch.run().cpu().numpy()) == _hex(oracle.reconstruct_exact(results, coefficients, observables))


@pytest.mark.parametrize("rounding", ["exact", "reference"])
def test_layout_cache_of_small_calls(lib, rounding):
    """Repeated small calls of one layout reuse plan, schedule and uploaded tables and stage their coefficients behind the
    shot bytes; new data, new coefficients and a changed layout must all give the oracle's result."""
    engine._skeletons.clear()
    want_fn = oracle.reconstruct_literal if rounding == "reference" else oracle.reconstruct_exact
    observables = _cases.SMALL_CASES["molecular"]()[2]
    paths = []
    for seed in (1, 2, 3):
        rng = np.random.default_rng(seed)
        coefficients = _cases.random_coefficients(rng, 5)
        results = _cases.make_results(observables, 5, 96, 3, seed=seed)  # same shapes, new bytes
        got = reconstruct_expectation_values(results, coefficients, observables, rounding=rounding)
        paths.append(engine.last_stage_trace["path"])
        assert _hex(got) == _hex(want_fn(results, coefficients, observables)), seed
    assert paths == ["simple", "simple (cached layout)", "simple (cached layout)"]
    # another shot count -> another layout; then the first layout again -> still cached
    results = _cases.make_results(observables, 5, 97, 3, seed=4)
    got = reconstruct_expectation_values(results, coefficients, observables, rounding=rounding)
    assert engine.last_stage_trace["path"] == "simple" and _hex(got) == _hex(want_fn(results, coefficients, observables))
    results = _cases.make_results(observables, 5, 96, 3, seed=5)
    got = reconstruct_expectation_values(results, coefficients, observables, rounding=rounding)
    assert engine.last_stage_trace["path"] == "simple (cached layout)"
    assert _hex(got) == _hex(want_fn(results, coefficients, observables))
    # shared result objects (duplicates) bypass the cached layout and are still merged
    dup = _with_duplicates(results, coefficients, [("S1", 0, 3)])
    got = reconstruct_expectation_values(dup, coefficients, observables, rounding=rounding)
    assert engine.last_stage_trace["path"] == "simple" and engine.last_stage_trace["dedup"]["dedup_ratio"] > 1.0
    assert _hex(got) == _hex(want_fn(dup, coefficients, observables))


def test_mid_size_calls_pipeline_gather_and_copy(lib):
    """4..64 MiB: the simple path hands the gather + copy to qcut_stage_h2d (2 MiB pieces, four threads)."""
    rng = np.random.default_rng(31)
    observables = {"A": _cases.random_paulis(rng, 6, 40, 1, 5, "Z")}
    coefficients = _cases.random_coefficients(rng, 3)
    results = _cases.make_results(observables, 3, 400_000, 2, seed=32)  # 3 x 400 000 x (5 + 1) bytes = 7.2 MB
    for _ in range(2):  # second call: cached layout, same pipeline
        got = reconstruct_expectation_values(results, coefficients, observables)
        assert _hex(got) == _hex(oracle.reconstruct_exact(results, coefficients, observables))
    assert engine.last_stage_trace["path"] == "simple (cached layout)"


@pytest.mark.parametrize("rounding", ["exact", "reference"])
def test_mixed_sampler_v1_and_v2_subsystems(lib, rounding):
    """The reference decides V1 / V2 per subsystem (cutting_reconstruction.py:143,150): SamplerV2 subsystems go through
    ingest + K1, SamplerV1 ones through the quasi kernel, one K2 launch consumes both in the caller's subsystem order."""
    results, coefficients, observables = _cases.case_mixed_samplers()
    want = oracle.reconstruct_literal if rounding == "reference" else oracle.reconstruct_exact
    got = reconstruct_expectation_values(results, coefficients, observables, rounding=rounding)
    assert _hex(got) == _hex(want(results, coefficients, observables))
    # the other way round, and with shared (duplicate) SamplerV2 PUBs
    flipped = {k: results[k] for k in ("B", "C", "A")}
    obs_flipped = {k: observables[k] for k in ("B", "C", "A")}
    got = reconstruct_expectation_values(flipped, coefficients, obs_flipped, rounding=rounding)
    assert _hex(got) == _hex(want(flipped, coefficients, obs_flipped))
    dup = _with_duplicates(results, coefficients, [("A", 0, 2), ("C", 1, 4)])
    got = reconstruct_expectation_values(dup, coefficients, observables, rounding=rounding)
    assert _hex(got) == _hex(want(dup, coefficients, observables))
