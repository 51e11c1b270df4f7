"""Sharding of a reconstruction over the GPUs of one node, one process per GPU.

Every PUB is independent in K1 and the product over subsystems only couples PUBs of the same
coefficient index ``i`` (reference ``cutting_reconstruction.py:134-168``).  Two ways to cut the work:

* **coefficient slices** (default): rank ``r`` takes a contiguous slice of ``i`` *for every subsystem*,
  balanced by shot bytes (shots may differ per PUB, reference ``:155``), computes its partial ``out[k]``
  locally; the only communication of the whole path is one all-reduce of ``K`` doubles (<= 8 KB).
* **tally table** (few coefficients: ``C < world`` or no balanced contiguous split exists, e.g. 36 tuples
  on 8 GPUs): the flattened (subsystem, PUB, shot range) list is cut into ``world`` pieces of equal
  bytes at tile granularity -- tallies are additive, so a PUB may be split between ranks -- every rank
  tallies its piece, the int64 tally table is all-reduced (exact, order independent) and every rank runs
  the product/sum kernel on the complete table: all ranks return bit-identical results.

NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests of the host logic.  A rank that fails still takes
part in the collective with an error flag, so all ranks raise together instead of hanging.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

# all-reducing the tally table is only worth it while the table is small (int64 entries)
MAX_TALLY_TABLE = 1 << 22
# coefficient slices are considered balanced up to this max/mean load
MAX_IMBALANCE = 1.05
# up to this many PUBs every rank looks at every PUB's shape to balance by bytes; above it a sample decides
FULL_WALK_PUBS = 1 << 16


class ShardedResult:
    """A ``PrimitiveResult``-like sequence of which this process holds only PUBs ``[lo, lo + len(pubs))``.

    ``len()`` is the global number of PUBs (what the reference's count check at ``:120-131`` sees); indexing
    an absent PUB raises ``IndexError``.  For jobs whose results are loaded shard by shard: with equal-count
    coefficient slices a rank only ever touches its own slice.
    """

    def __init__(self, total: int, lo: int, pubs):
        self.total, self.lo, self.pubs = int(total), int(lo), pubs

    @property
    def resident_range(self) -> tuple[int, int]:
        return self.lo, self.lo + len(self.pubs)

    def __len__(self) -> int:
        return self.total

    def __getitem__(self, idx):
        j = idx - self.lo
        if not 0 <= j < len(self.pubs):
            raise IndexError(f"PUB {idx} is not resident in this process (it holds {self.lo}..{self.lo + len(self.pubs) - 1})")
        return self.pubs[j]


def rank_and_world(process_group=None) -> tuple[int, int]:
    """Sharding is opt-in: without an explicit process group the call is single-process (drop-in behaviour)."""
    if process_group is None or not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(process_group), dist.get_world_size(process_group)


def coefficient_slice(n_coeffs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, count-balanced slice ``[lo, hi)`` of the coefficient indices owned by ``rank``."""
    base, extra = divmod(n_coeffs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_slices(weights: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous slices of ``range(len(weights))`` whose weight sums are as equal as contiguity allows:
    boundary ``r`` sits where the running sum is closest to ``r / world`` of the total (integer arithmetic)."""
    w = np.asarray(weights, dtype=np.int64)
    n = len(w)
    cum = np.concatenate([[0], np.cumsum(w)])
    total = int(cum[-1])
    if total == 0:
        return [coefficient_slice(n, r, world) for r in range(world)]
    bounds = [0]
    for r in range(1, world):
        target = total * r  # compare world * cum with total * r: no division, no rounding
        j = int(np.searchsorted(cum * world, target, side="left"))
        if j > 0 and abs(int(cum[j - 1]) * world - target) <= abs(int(cum[min(j, n)]) * world - target):
            j -= 1
        bounds.append(min(max(j, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def imbalance(weights: np.ndarray, slices: list[tuple[int, int]]) -> float:
    """max / mean load of the slices (1.0 = perfectly balanced)."""
    cum = np.concatenate([[0], np.cumsum(np.asarray(weights, dtype=np.int64))])
    loads = np.array([int(cum[hi] - cum[lo]) for lo, hi in slices], dtype=np.float64)
    mean = loads.mean()
    return float(loads.max() / mean) if mean > 0 else 1.0


def piece_ranges(row_bytes: np.ndarray, shots: np.ndarray, world: int, tile: int) -> list[list[tuple[int, int, int]]]:
    """Cut the flattened PUB list into ``world`` byte-balanced pieces at tile granularity.

    Returns, per rank, ``(pub, shot_lo, shot_hi)`` triples.  A boundary that falls inside PUB ``p`` is rounded
    to a multiple of ``tile`` shots by one rule shared by both neighbours, so the pieces partition every PUB.
    """
    row_bytes = np.asarray(row_bytes, dtype=np.int64)
    shots = np.asarray(shots, dtype=np.int64)
    size = row_bytes * shots
    start = np.concatenate([[0], np.cumsum(size)])
    total = int(start[-1])
    n = len(size)

    def boundary(r: int) -> tuple[int, int]:  # -> (pub, shot) position of cut r
        if r <= 0:
            return 0, 0
        if r >= world or total == 0:
            return n, 0
        x = total * r // world
        p = int(np.searchsorted(start, x, side="right")) - 1
        p = min(max(p, 0), n - 1)
        if row_bytes[p] == 0 or shots[p] == 0:
            return p, 0
        s = (x - int(start[p])) // int(row_bytes[p])
        s = min(int(shots[p]), (s + tile // 2) // tile * tile)
        return (p, s) if s < shots[p] else (p + 1, 0)

    out = []
    for r in range(world):
        (p0, s0), (p1, s1) = boundary(r), boundary(r + 1)
        pieces = []
        for p in range(p0, min(p1 + 1, n)):
            lo = s0 if p == p0 else 0
            hi = s1 if p == p1 else int(shots[p])
            if hi > lo:
                pieces.append((p, lo, hi))
        out.append(pieces)
    return out


@dataclass
class Sharding:
    """How one call is cut over the ranks (identical decision on every rank)."""

    mode: str  # "coefficients" | "tallies"
    rank: int = 0
    world: int = 1
    i_lo: int = 0
    i_hi: int = 0
    balanced_by: str = "count"  # "count" | "bytes"
    imbalance: float = 1.0
    collected: dict = field(default_factory=dict)  # label index -> walked arrays of ALL its PUBs (tally mode)


def all_reduce_sum(partial: torch.Tensor, process_group=None, world: int | None = None) -> torch.Tensor:
    """Sum the per-rank partial expectation values (issued on the compute stream for NCCL)."""
    if world is None:
        world = rank_and_world(process_group)[1]
    if world > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=process_group)
    return partial


class RemoteRankError(RuntimeError):
    """Another rank of the process group failed inside the same reconstruction call."""


def all_reduce_with_flag(payload: torch.Tensor, failed: bool, process_group, world: int) -> torch.Tensor:
    """All-reduce ``payload`` (1-D; its LAST element is reserved for the error flag; a failed rank passes zeros
    elsewhere).  Nothing synchronises here: the caller reads the flag with the result (:func:`check_flag`)."""
    if world > 1:
        payload[-1] = 1 if failed else 0
        dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=process_group)
    return payload


def agree_on_failure(failed: bool, process_group, world: int) -> bool:
    """True on every rank if any rank reports a failure (an all-reduce of one flag; synchronises with the host)."""
    if world <= 1:
        return failed
    backend = dist.get_backend(process_group)
    flag = torch.tensor([1 if failed else 0], dtype=torch.int32, device="cuda" if backend == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=process_group)
    return bool(flag.item())


def check_flag(flag_value) -> None:
    if float(flag_value) != 0:
        raise RemoteRankError("reconstruct_expectation_values failed on at least one rank of the process group")
