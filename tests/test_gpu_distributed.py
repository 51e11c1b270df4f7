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
for name in ("molecular", "heavy_hex", "tutorial"):
    results, coefficients, observables = _cases.SMALL_CASES[name]()
    got = np.array(reconstruct_expectation_values(results, coefficients, observables, process_group=dist.group.WORLD))
    want = np.array(oracle.reconstruct_exact(results, coefficients, observables))
    kappa = max(1.0, sum(abs(c) for c, _ in coefficients))
    assert np.max(np.abs(got - want)) <= 1e-13 * kappa, (name, rank, got, want)
    gathered = [torch.zeros(len(got), dtype=torch.float64, device="cuda") for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, torch.tensor(got, device="cuda"))
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_reconstruction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, QCUT_ROOT=str(ROOT))
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29631", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert proc.stdout.count("ok") >= 2
