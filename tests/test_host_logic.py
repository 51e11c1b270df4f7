"""Host logic without a GPU: grouping invariants (property based), mask layout, install(), staged-size accounting."""

import sys
import types

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from qiskit_addon_cutting_b200 import PauliList, install, reconstruct_expectation_values
from qiskit_addon_cutting_b200.observable_grouping import CommutingObservableGroup, ObservableCollection, most_general_observable

labels = st.integers(1, 9).flatmap(
    lambda n: st.lists(st.text(alphabet="IXYZ", min_size=n, max_size=n), min_size=1, max_size=14)
)


@settings(max_examples=120, deadline=None)
@given(labels)
def test_grouping_invariants(obs_labels):
    paulis = PauliList(obs_labels)
    coll = ObservableCollection(paulis)
    seen = {}
    for gi, cog in enumerate(coll.groups):
        members = cog.commuting_observables
        # every member is measurable from the general observable: same Pauli wherever the member is non-identity
        gz, gx = np.asarray(cog.general_observable.z), np.asarray(cog.general_observable.x)
        for ti, m in enumerate(members):
            non_i = np.asarray(m.z) | np.asarray(m.x)
            assert np.all(np.asarray(m.z)[non_i] == gz[non_i]) and np.all(np.asarray(m.x)[non_i] == gx[non_i])
            assert m.to_label() not in seen, "an observable sits in exactly one group"
            seen[m.to_label()] = (gi, ti)
            # mask bit i <-> qubit pauli_indices[i]
            expect = sum(1 << i for i, q in enumerate(cog.pauli_indices) if non_i[q])
            assert cog.pauli_bitmasks[ti] == expect
        assert cog.pauli_indices == sorted(cog.pauli_indices)
        assert set(cog.pauli_indices) == set(np.flatnonzero(gz | gx).tolist())
    assert set(seen) == set(obs_labels)
    for k, lab in enumerate(obs_labels):
        assert coll.lookup_indices[k] == [seen[lab]]
    # the oracle's pure-Python grouping (reference loops) agrees with the vectorised one
    oc = oracle.build_collection(paulis)
    assert [g["masks"] for g in oc["groups"]] == [c.pauli_bitmasks for c in coll.groups]
    assert [g["pauli_indices"] for g in oc["groups"]] == [c.pauli_indices for c in coll.groups]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 70), st.integers(0, 2**70 - 1), st.integers(1, 12))
def test_mask_bytes_are_big_endian_rows(nbits, value, nbytes):
    cog = CommutingObservableGroup.__new__(CommutingObservableGroup)
    object.__setattr__(cog, "pauli_bitmasks", [value & ((1 << nbits) - 1)])
    rows = cog.mask_bytes(nbytes)
    assert rows.shape == (1, nbytes)
    assert int.from_bytes(rows[0].tobytes(), "big") == (value & ((1 << nbits) - 1)) & ((1 << (8 * nbytes)) - 1)


def test_most_general_observable_matches_reference_docstring_and_errors():
    assert most_general_observable(PauliList(["IIIZ", "IIZZ", "XIII"])).to_label() == "XIZZ"  # observable_grouping.py:77-78
    with pytest.raises(ValueError, match="Empty input sequence"):
        most_general_observable([])
    with pytest.raises(ValueError, match="Observables are incompatible"):
        most_general_observable(PauliList(["X", "Z"]))
    with pytest.raises(ValueError, match="incorrect qubit count"):
        most_general_observable(list(PauliList(["XX"])) + list(PauliList(["X"])))
    with pytest.raises(ValueError, match="only supports Paulis with phase == 0"):
        CommutingObservableGroup(PauliList(["ZZ"])[0], list(PauliList(["iZZ"])))


def test_install_patches_the_reference_when_importable(monkeypatch):
    assert install() is False  # the reference is not installed in this image
    pkg = types.ModuleType("qiskit_addon_cutting")
    sub = types.ModuleType("qiskit_addon_cutting.cutting_reconstruction")
    sub.reconstruct_expectation_values = pkg.reconstruct_expectation_values = lambda *a, **k: "reference"
    pkg.cutting_reconstruction = sub
    monkeypatch.setitem(sys.modules, "qiskit_addon_cutting", pkg)
    monkeypatch.setitem(sys.modules, "qiskit_addon_cutting.cutting_reconstruction", sub)
    assert install() is True
    assert pkg.reconstruct_expectation_values is reconstruct_expectation_values
    assert sub.reconstruct_expectation_values is reconstruct_expectation_values


def test_compiled_kernels_are_cached_on_disk(lib, tmp_path, monkeypatch):
    """Cubins are stored under $QCUT_CACHE_DIR keyed by the source; damaged files are replaced; "" disables."""
    import time

    from qiskit_addon_cutting_b200 import _lib as L

    groups = np.array([(14, 2, 0, 0)], dtype=L.GROUP_DTYPE)
    masks = np.random.default_rng(5).integers(0, 256, size=28, dtype=np.uint8)
    cache = tmp_path / "cubins"
    monkeypatch.setenv("QCUT_CACHE_DIR", str(cache))
    t0 = time.perf_counter()
    n, size = L.plan_compile_check(groups, masks)
    cold = time.perf_counter() - t0
    files = sorted(cache.glob("*.cubin"))
    assert n == 1 and len(files) == 1 and files[0].stat().st_size == size and files[0].read_bytes()[:4] == b"\x7fELF"
    t0 = time.perf_counter()
    assert L.plan_compile_check(groups, masks) == (n, size)
    assert time.perf_counter() - t0 < cold / 3  # no NVRTC run
    assert not list(cache.glob("*.tmp"))
    # a different mask set is a different kernel
    masks2 = masks.copy()
    masks2[0] ^= 1
    L.plan_compile_check(groups, masks2)
    assert len(list(cache.glob("*.cubin"))) == 2
    # a damaged file is not trusted
    files[0].write_bytes(b"garbage")
    assert L.plan_compile_check(groups, masks) == (n, size)
    assert files[0].read_bytes()[:4] == b"\x7fELF"
    # switched off
    monkeypatch.setenv("QCUT_CACHE_DIR", "")
    before = len(list(cache.glob("*")))
    assert L.plan_compile_check(groups, masks) == (n, size)
    assert len(list(cache.glob("*"))) == before


def _sets(rng, n_sets, n_terms, nb):
    """n_sets distinct groups of n_terms random masks on nb-byte rows: (group records, mask bytes) for plan creation."""
    from qiskit_addon_cutting_b200 import _lib as L

    groups = np.zeros(n_sets, dtype=L.GROUP_DTYPE)
    chunks, offset = [], 0
    for g in range(n_sets):
        rows = rng.integers(0, 256, size=(n_terms, nb), dtype=np.uint8)
        rows[:, 0] |= 1  # no all-zero row, sets differ with overwhelming probability
        groups[g] = (n_terms, nb, offset, 0)
        chunks.append(rows.reshape(-1))
        offset += rows.size
    return groups, np.concatenate(chunks)


def _merge(parts):
    from qiskit_addon_cutting_b200 import _lib as L

    groups = np.concatenate([g for g, _ in parts]).astype(L.GROUP_DTYPE)
    masks, offset, first = [], 0, 0
    for g, m in parts:
        groups["mask_offset"][first:first + len(g)] = g["mask_offset"] + offset
        masks.append(m)
        offset += m.size
        first += len(g)
    return groups, np.concatenate(masks)


def test_which_groups_share_a_kernel(lib, tmp_path, monkeypatch):
    """Kernel assignment of a plan (api.cu: assign_jit_kernels), read off the number of NVRTC kernels it compiles: groups of
    more than 32 terms get a kernel each; smaller ones share a switch kernel per row width (64 sets each) when a width has
    at least 4 sets or is alone, a mixed-width kernel (24 sets each) otherwise."""
    from qiskit_addon_cutting_b200 import _lib as L

    monkeypatch.setenv("QCUT_CACHE_DIR", str(tmp_path))
    rng = np.random.default_rng(11)
    count = lambda parts: L.plan_compile_check(*_merge(parts))[0]  # noqa: E731
    assert count([_sets(rng, 2, 33, 2)]) == 2                      # two big groups: a kernel each
    assert count([_sets(rng, 70, 5, 2)]) == 2                      # one width, 70 small sets: 64 + 6 behind switches
    assert count([_sets(rng, 5, 5, 1), _sets(rng, 6, 7, 2)]) == 2  # two widths with >= 4 sets each: one switch kernel per width
    assert count([_sets(rng, 2, 5, 1), _sets(rng, 3, 7, 3), _sets(rng, 1, 9, 6)]) == 1  # a handful per width: one mixed kernel
    assert count([_sets(rng, 1, 40, 8), _sets(rng, 3, 6, 8)]) == 2  # a big group and three small ones of the only small width
    monkeypatch.setenv("QCUT_JIT_SWITCH", "0")
    assert count([_sets(rng, 70, 5, 2)]) == 3                      # the older scheme: 24 functions per kernel
