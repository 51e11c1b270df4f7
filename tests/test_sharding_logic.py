"""Host logic of the multi-GPU sharding (CPU only): byte-balanced coefficient slices, the tally-table pieces
(fallback for C < R), the decision between the two, and the error-flag collective (gloo)."""

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _cases
from oracle import oracle
from qiskit_addon_cutting_b200 import distributed as D
from qiskit_addon_cutting_b200 import engine
from qiskit_addon_cutting_b200.containers import PauliList, PrimitiveResult
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection

ROOT = Path(__file__).resolve().parents[1]


def test_balanced_slices_cover_and_beat_equal_count_on_ragged_weights():
    rng = np.random.default_rng(1)
    for n, world in [(0, 4), (1, 4), (5, 8), (97, 8), (1000, 3), (36, 8)]:
        weights = rng.integers(1, 1000, size=n) * rng.choice([1, 1, 50], size=n)
        slices = D.balanced_slices(weights, world)
        assert len(slices) == world and slices[0][0] == 0 and slices[-1][1] == n
        assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(slices, slices[1:]))
        if n >= 4 * world:
            equal = [D.coefficient_slice(n, r, world) for r in range(world)]
            assert D.imbalance(weights, slices) <= D.imbalance(weights, equal) + 1e-12
    # uniform weights -> the equal-count split
    assert D.balanced_slices(np.full(16, 7), 4) == [(0, 4), (4, 8), (8, 12), (12, 16)]
    # a step profile: 8 heavy tuples then 56 light ones
    w = np.array([100] * 8 + [1] * 56)
    assert D.imbalance(w, D.balanced_slices(w, 4)) < D.imbalance(w, [D.coefficient_slice(64, r, 4) for r in range(4)])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_piece_ranges_partition_every_pub(world):
    rng = np.random.default_rng(world)
    for trial in range(20):
        n = int(rng.integers(1, 40))
        shots = rng.integers(0, 60_000, size=n)
        row = rng.integers(2, 12, size=n)
        tile = 4096
        ranks = D.piece_ranges(row, shots, world, tile)
        assert len(ranks) == world
        covered = np.zeros(n, dtype=np.int64)
        last = (-1, 0)
        for pieces in ranks:
            for p, lo, hi in pieces:
                assert 0 <= lo < hi <= shots[p]
                assert lo % tile == 0 and (hi % tile == 0 or hi == shots[p])
                assert (p, lo) >= last  # pieces come in order and never overlap
                assert lo == covered[p]
                covered[p] = hi
                last = (p, hi)
        assert np.array_equal(covered, shots)
        total = int((row * shots).sum())
        loads = [sum((hi - lo) * int(row[p]) for p, lo, hi in pieces) for pieces in ranks]
        if total:
            assert max(loads) <= total / world + 2 * tile * int(row.max())


def _ragged_case(num_coeffs=12, seed=5):
    observables = {"A": PauliList(["ZI", "IZ", "XX"]), "B": PauliList(["ZZ", "XI", "IX"])}
    # shots fall with the coefficient index, as when shots are allocated by |coefficient|
    results = _cases.make_results(observables, num_coeffs, lambda label, i, g: 4000 // (i + 1) + 7, 2, seed)
    coeffs = _cases.random_coefficients(np.random.default_rng(seed), num_coeffs)
    return results, coeffs, observables


def _plan(results, observables, n_coeffs, world, rank=0, **kw):
    colls = [ObservableCollection(o) for o in observables.values()]
    return engine.plan_sharding([results[k] for k in observables], colls, n_coeffs, rank, world, **kw)


def test_plan_sharding_decisions():
    results, coeffs, observables = _cases.case_tutorial()  # 36 tuples, equal shots
    s = _plan(results, observables, len(coeffs), 4)
    assert (s.mode, s.i_lo, s.i_hi) == ("coefficients", 0, 9)  # 36 / 4 is balanced
    s = _plan(results, observables, len(coeffs), 8, rank=7)
    assert s.mode == "tallies"  # 5/5/5/5/4/4/4/4 tuples: 11 % imbalance -> split the PUBs by bytes instead
    assert _plan(results, observables, len(coeffs), 8, allow_tally_mode=False).mode == "coefficients"
    results, coeffs, observables = _cases.case_heavy_hex()
    if len(coeffs) < 8:
        assert _plan(results, observables, len(coeffs), 8).mode == "tallies"  # C < R
    # ragged shots: slices are balanced by bytes, identically on every rank
    results, coeffs, observables = _ragged_case(64)
    plans = [_plan(results, observables, 64, 4, rank=r, allow_tally_mode=False) for r in range(4)]
    assert all(p.balanced_by == "bytes" for p in plans)
    edges = [(p.i_lo, p.i_hi) for p in plans]
    assert edges[0][0] == 0 and edges[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert edges[0][1] - edges[0][0] < 16 < edges[-1][1] - edges[-1][0]  # heavy tuples first -> fewer of them
    # sharded inputs: nothing to look at, equal-count slices
    G = len(ObservableCollection(observables["A"]).groups)
    sharded = {k: D.ShardedResult(len(v), 16 * (len(v) // 64), [v[i] for i in range(16 * (len(v) // 64), 32 * (len(v) // 64))])
               for k, v in results.items()}
    s = _plan(sharded, observables, 64, 4, rank=1)
    assert (s.mode, s.i_lo, s.i_hi, G > 0) == ("coefficients", 16, 32, True)


def test_sharded_result_protocol():
    r = D.ShardedResult(10, 4, ["a", "b", "c"])
    assert len(r) == 10 and r[4] == "a" and r[6] == "c" and r.resident_range == (4, 7)
    with pytest.raises(IndexError):
        r[3]
    with pytest.raises(IndexError):
        r[7]


def _tally_table_worker(rank, world, port, out_path, fail_rank):
    for p in (ROOT, ROOT / "tests"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    results, coefficients, observables = _ragged_case(2)  # C = 2 < world = 3
    labels = list(observables)
    colls = [ObservableCollection(observables[k]) for k in labels]
    sharding = engine.plan_sharding([results[k] for k in labels], colls, len(coefficients), rank, world)
    assert sharding.mode == "tallies"
    coeffs = np.array([c for c, _ in coefficients])
    tables, arrays = engine.build_host_tables([results[k] for k in labels], colls, coeffs, 0, len(coeffs), 3,
                                              collected=sharding.collected)
    nb = tables.groups["nb_obs"][tables.pubs["group"]].astype(np.int64)
    pieces = D.piece_ranges(nb + tables.pubs["nb_qpd"], tables.pubs["shots"], world, 512)[rank]
    # this rank's pieces, tallied by the oracle (the GPU version of the same flow: tests/test_gpu_distributed.py)
    flat = [pub for k in labels for pub in results[k]]
    table = torch.zeros(tables.n_tallies + 1, dtype=torch.int64)
    for p, lo, hi in pieces:
        pub, grp = tables.pubs[p], tables.groups[tables.pubs[p]["group"]]
        T, w = int(grp["n_terms"]), int(grp["nb_obs"])
        masks = [int.from_bytes(tables.masks[int(grp["mask_offset"]) + t * w: int(grp["mask_offset"]) + (t + 1) * w].tobytes(), "big")
                 for t in range(T)]
        o = flat[p].data.observable_measurements.array[lo:hi]
        q = flat[p].data.qpd_measurements.array[lo:hi]
        table[int(pub["tally_offset"]): int(pub["tally_offset"]) + T] += torch.from_numpy(oracle.tally_exact(o, q, masks))
    D.all_reduce_with_flag(table, rank == fail_rank, dist.group.WORLD, world)
    if rank == 0:
        np.save(out_path, table.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 2])
def test_tally_table_all_reduce_with_more_ranks_than_coefficients(tmp_path, fail_rank):
    out = tmp_path / "table.npy"
    port = 23000 + os.getpid() % 2000 + (7 if fail_rank >= 0 else 0)
    mp.spawn(_tally_table_worker, args=(3, port, str(out), fail_rank), nprocs=3, join=True)
    table = np.load(out)
    results, coefficients, observables = _ragged_case(2)
    want = np.concatenate([t for k in observables for t in oracle.all_tallies(results, coefficients, observables)[k]])
    assert table[-1] == (1 if fail_rank >= 0 else 0)  # the error flag reaches every rank
    assert np.array_equal(table[:-1], want)
    if fail_rank >= 0:
        with pytest.raises(D.RemoteRankError):
            D.check_flag(table[-1])


def test_slice_tables_for_strong_scaling():
    import workloads as W

    w = W.get_workload("cfg5_molecular", 0.2)
    tables = W.build_tables(w)
    half = W.slice_tables(tables, 0, w.n_coeffs // 2)
    assert len(half.coeffs) == w.n_coeffs // 2
    assert half.n_tallies == int(tables.groups["n_terms"][half.pubs["group"]].sum())
    ends = [int(lab["pub_base"]) for lab in half.labels[1:]] + [len(half.pubs)]
    for lab, lab_full, end in zip(half.labels, tables.labels, ends):
        G = int(lab["num_groups"])
        assert end - int(lab["pub_base"]) == (w.n_coeffs // 2) * G
        a = half.pubs[int(lab["pub_base"]): end]
        b = tables.pubs[int(lab_full["pub_base"]): int(lab_full["pub_base"]) + len(a)]
        assert np.array_equal(a["obs_offset"], b["obs_offset"]) and np.array_equal(a["group"], b["group"])
    assert half.work * 2 <= tables.work + tables.work // w.n_coeffs
