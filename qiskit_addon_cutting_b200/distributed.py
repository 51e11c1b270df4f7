"""Sharding of the subexperiment tuples over the GPUs of one node, one process per GPU.

Every PUB is independent in K1 and the product over subsystems only couples PUBs of the same
coefficient index ``i`` (reference ``cutting_reconstruction.py:134-168``), so rank ``r`` takes a
contiguous slice of ``i`` *for every subsystem*, computes its partial ``out[k]`` locally and the
only communication of the whole path is one all-reduce of ``K`` doubles (<= 8 KB): NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests of the host logic.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def rank_and_world(process_group=None) -> tuple[int, int]:
    """Sharding is opt-in: without an explicit process group the call is single-process (drop-in behaviour)."""
    if process_group is None or not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(process_group), dist.get_world_size(process_group)


def coefficient_slice(n_coeffs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice ``[lo, hi)`` of the coefficient indices owned by ``rank``."""
    base, extra = divmod(n_coeffs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum(partial: torch.Tensor, process_group=None, world: int | None = None) -> torch.Tensor:
    """Sum the per-rank partial expectation values (issued on the compute stream for NCCL)."""
    if world is None:
        world = rank_and_world(process_group)[1]
    if world > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=process_group)
    return partial
