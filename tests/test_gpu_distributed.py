"""Two ranks, two GPUs, NCCL: the sharded API call returns on every rank what a single process returns."""

import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["QCUT_ROOT"]); sys.path.insert(0, os.path.join(os.environ["QCUT_ROOT"], "tests"))
import _cases
from oracle import oracle
from qiskit_addon_cutting_b200 import reconstruct_expectation_values
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from qiskit_addon_cutting_b200 import engine, distributed
from qiskit_addon_cutting_b200.observable_grouping import ObservableCollection
from qiskit_addon_cutting_b200.containers import PrimitiveResult, SamplerResult, QuasiDistribution, PauliList, WeightType
world = dist.get_world_size()

def check(name, results, coefficients, observables, expect_mode=None, **kw):
    if expect_mode is not None:
        labels = list(observables)
        colls = [ObservableCollection(observables[k]) for k in labels]
        mode = engine.plan_sharding([results[k] for k in labels], colls, len(coefficients), rank, world).mode
        assert mode == expect_mode, (name, mode)
    got = np.array(reconstruct_expectation_values(results, coefficients, observables, process_group=dist.group.WORLD, **kw))
    want = np.array(oracle.reconstruct_exact(results, coefficients, observables))
    kappa = max(1.0, sum(abs(c) for c, _ in coefficients))
    assert np.max(np.abs(got - want)) <= 1e-13 * kappa, (name, rank, got, want)
    gathered = [torch.zeros(len(got), dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(got, device="cuda"))
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree"
    return got

for name in ("molecular", "heavy_hex", "tutorial"):
    check(name, *_cases.SMALL_CASES[name]())
# C < world (SURVEY 8e fallback): the ranks split the PUBs by bytes and all-reduce the int64 tally table
results, coefficients, observables = _cases.SMALL_CASES["heavy_hex"]()
one = {k: PrimitiveResult(list(v)[: len(v) // len(coefficients)]) for k, v in results.items()}
got = check("one tuple", one, coefficients[:1], observables, expect_mode="tallies")
single = np.array(reconstruct_expectation_values(one, coefficients[:1], observables))
assert np.array_equal(got, single), "tally-table sharding must be bit-identical to a single process"
# an unbalanced count (5 tuples on 2 ranks) also takes the tally-table route; so do ragged shot counts
check("molecular/tallies", *_cases.SMALL_CASES["molecular"](), expect_mode="tallies" if world == 2 else None)
# reference rounding cannot split a PUB (its chains are sequential): coefficient slices, some of them empty
check("one tuple, reference rounding", one, coefficients[:1], observables, rounding="reference")
# SamplerV1 inputs with more ranks than tuples: ranks with an empty slice contribute zeros and do not hang
v1 = {"A": SamplerResult([QuasiDistribution({0: 0.5, 3: 0.25, 5: 0.25})]), "B": SamplerResult([QuasiDistribution({1: 0.75, 2: 0.25})])}
obs_v1 = {"A": PauliList(["ZZ"]), "B": PauliList(["ZI"])}
got = reconstruct_expectation_values(v1, [(0.5, WeightType.EXACT)], obs_v1, process_group=dist.group.WORLD)
assert got == reconstruct_expectation_values(v1, [(0.5, WeightType.EXACT)], obs_v1), got
# a failure on ONE rank (a PUB with fewer observable rows than qpd rows) raises on EVERY rank instead of hanging
bad = dict(results)
if rank == world - 1:
    pubs = list(bad["B"])
    broken = pubs[-1].data
    class Reg:
        def __init__(self, array): self.array = array
    class Data: pass
    d = Data(); d.observable_measurements = Reg(broken.observable_measurements.array[:5]); d.qpd_measurements = broken.qpd_measurements
    class Pub: pass
    pb = Pub(); pb.data = d
    pubs[-1] = pb
    bad["B"] = PrimitiveResult(pubs)
try:
    reconstruct_expectation_values(bad, coefficients, observables, process_group=dist.group.WORLD)
    raise SystemExit("expected an exception on every rank")
except (IndexError, distributed.RemoteRankError) as exc:
    assert isinstance(exc, IndexError) == (rank == world - 1), (rank, type(exc))
check("after the failure", results, coefficients, observables)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("nproc", [2, 4])
def test_two_rank_nccl_reconstruction(tmp_path, nproc):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, QCUT_ROOT=str(ROOT))
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
         "--master-port", "29631", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert proc.stdout.count("ok") >= nproc
