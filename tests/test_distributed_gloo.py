"""World-size-2 check of the sharding logic on CPU (gloo): contiguous coefficient slices of every subsystem,
per-rank partial sums, one all-reduce.  The per-rank compute is stood in by the oracle (no GPU here); the GPU
version of the same flow is exercised by bench.py --gpus N and tests/test_gpu_distributed.py."""

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out_path):
    for p in (ROOT, ROOT / "tests"):
        sys.path.insert(0, str(p))
    import _cases
    from oracle import oracle
    from qiskit_addon_cutting_b200 import distributed
    from qiskit_addon_cutting_b200.containers import PrimitiveResult

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    results, coefficients, observables = _cases.case_molecular()
    assert distributed.rank_and_world(None) == (0, 1), "sharding must be opt-in"
    r, w = distributed.rank_and_world(dist.group.WORLD)
    lo, hi = distributed.coefficient_slice(len(coefficients), r, w)
    # this rank's PUBs: result[i * G + g] for i in [lo, hi), for EVERY subsystem
    shard = {}
    for label, res in results.items():
        G = len(res) // len(coefficients)
        shard[label] = PrimitiveResult([res[i * G + g] for i in range(lo, hi) for g in range(G)])
    partial = torch.tensor(oracle.reconstruct_exact(shard, coefficients[lo:hi], observables), dtype=torch.float64)
    total = distributed.all_reduce_sum(partial, dist.group.WORLD, w)
    if rank == 0:
        np.save(out_path, total.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_coefficient_slices_cover_everything():
    from qiskit_addon_cutting_b200.distributed import coefficient_slice

    for n in (0, 1, 5, 8, 97000):
        for world in (1, 2, 3, 8):
            edges = [coefficient_slice(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1


def test_two_rank_sharded_reconstruction_matches_single_process(tmp_path):
    import _cases
    from oracle import oracle

    out = tmp_path / "out.npy"
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    results, coefficients, observables = _cases.case_molecular()
    want = np.array(oracle.reconstruct_exact(results, coefficients, observables))
    got = np.load(out)
    assert np.max(np.abs(got - want)) <= 1e-14 * max(1.0, sum(abs(c) for c, _ in coefficients))
