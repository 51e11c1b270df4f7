"""Host-side engine: plan tables, staging into HBM, and launches through the C ABI.

PyTorch is used for device memory, pinned host memory, streams and (in ``distributed.py``)
NCCL -- plumbing only.  All arithmetic of the path runs in the kernels of
``libqcut_b200.so``; nothing here falls back to the CPU.

Terminology follows the reference: a *PUB* is the result of one subexperiment
(``result[i * G + g]``, reference ``cutting_reconstruction.py:142``); a *group* is a
``CommutingObservableGroup``; a *term* is one Pauli of a group; a *label* is a subsystem.
"""

from __future__ import annotations

import ctypes as C
import os
import queue
import threading
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import GROUP_DTYPE, LABEL_DTYPE, PUB_DTYPE, SLOT_DTYPE, TILE_SHOTS, check

_ALIGN = 16


def _align(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _to_device(a: np.ndarray) -> torch.Tensor:
    """Copy a (possibly structured) numpy array to the current CUDA device as raw bytes."""
    return _upload_packed([a])[0]


def _upload_packed(arrays: list[np.ndarray]) -> list[torch.Tensor]:
    """Several (possibly structured) numpy tables -> ONE host-to-device copy; returns a byte view per table.

    A reconstruction call uploads four or five small descriptor tables; one packed copy instead of one
    synchronous copy each matters for the small configurations (tens of microseconds per copy)."""
    raws = [np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in arrays]
    starts, total = [], 0
    for r in raws:
        starts.append(total)
        total += max(_align(r.size), _ALIGN)
    packed = np.zeros(total, dtype=np.uint8)
    for r, o in zip(raws, starts):
        packed[o: o + r.size] = r
    dev = torch.from_numpy(packed).to("cuda", non_blocking=False)
    return [dev[o: o + max(_align(r.size), _ALIGN)] for r, o in zip(raws, starts)]


# ------------------------------------------------------------------------------------------------
# Plan: group table + masks in HBM, kernels specialised by NVRTC (cached per mask set)
# ------------------------------------------------------------------------------------------------
class Plan:
    def __init__(self, groups: np.ndarray, masks: np.ndarray, *, flags: int = 0):
        _lib.require_device()
        self.groups = np.ascontiguousarray(groups, dtype=GROUP_DTYPE)
        self.masks = np.ascontiguousarray(masks, dtype=np.uint8).reshape(-1)
        handle = C.c_void_p()
        check(
            _lib.load().qcut_plan_create(
                _lib.np_ptr(self.groups),
                len(self.groups),
                _lib.np_ptr(self.masks) if self.masks.size else None,
                self.masks.size,
                flags,
                C.byref(handle),
            ),
            "qcut_plan_create",
        )
        self.handle = handle
        self.device = torch.cuda.current_device()

    def kernel_of(self, group: int) -> int:
        return _lib.load().qcut_plan_group_kernel(self.handle, group)

    def jit_state(self) -> int:
        """0: no specialised kernels requested, 1: being compiled in the background, 2: ready, -1: failed."""
        return _lib.load().qcut_plan_jit_state(self.handle)

    def wait(self) -> "Plan":
        """Block until a background compilation (``PLAN_ASYNC_JIT``) has finished."""
        check(_lib.load().qcut_plan_wait(self.handle), "qcut_plan_wait")
        return self

    def kernels(self) -> list[int]:
        return [self.kernel_of(g) for g in range(len(self.groups))]

    def source(self, kernel: int) -> str:
        lib = _lib.load()
        need = lib.qcut_plan_source(self.handle, kernel, None, 0)
        if need == 0:
            return ""
        buf = C.create_string_buffer(need)
        lib.qcut_plan_source(self.handle, kernel, buf, need)
        return buf.value.decode()

    def __del__(self):
        handle, self.handle = getattr(self, "handle", None), None
        if handle:
            try:
                _lib.load().qcut_plan_destroy(handle)
            except Exception:  # interpreter shutdown
                pass


_plan_cache: dict = {}
_plan_lock = threading.Lock()


_SYNC_JIT = os.environ.get("QCUT_PLAN_SYNC", "0") not in ("0", "", "off")
# Calls that stage less than this are launch bound whatever kernel runs (cfg1: 11.6 us specialised vs 14.0 us with
# run-time masks; cfg2: 19.9 vs 24.8 us): they use the ahead-of-time kernels and never invoke NVRTC (a 1 s cold start).
_SMALL_CALL_BYTES = 64 << 20


def get_plan(groups: np.ndarray, masks: np.ndarray, *, flags: int = 0, background: bool = False) -> Plan:
    """The plan for these groups (cached per device and mask content).  ``background``: if the plan has to be
    created, its NVRTC specialisation runs on a background thread (``QCUT_PLAN_ASYNC_JIT``) and the call is served
    by the ahead-of-time kernels until it is ready -- no cold start; ``QCUT_PLAN_SYNC=1`` switches that off."""
    key = (torch.cuda.current_device(), flags, groups.tobytes(), masks.tobytes())
    with _plan_lock:
        plan = _plan_cache.get(key)
        if plan is None:
            if len(_plan_cache) >= 16:
                _plan_cache.pop(next(iter(_plan_cache)))
            create = flags | (_lib.PLAN_ASYNC_JIT if background and not _SYNC_JIT else 0)
            plan = _plan_cache[key] = Plan(groups, masks, flags=create)
        return plan


# ------------------------------------------------------------------------------------------------
# Problem description (host tables) for one call on one rank
# ------------------------------------------------------------------------------------------------
@dataclass
class HostTables:
    """Everything the kernels need besides the shot bytes, as numpy tables.

    ``pubs`` covers the PUBs of every label for the coefficient range owned by this rank,
    label after label (``labels[l].pub_base``), each label ordered ``i * G + g``.
    """

    groups: np.ndarray  # GROUP_DTYPE, one row per (label, group, nb_obs) variant
    masks: np.ndarray  # uint8
    pubs: np.ndarray  # PUB_DTYPE
    labels: np.ndarray  # LABEL_DTYPE
    lookup_start: np.ndarray  # uint32 [L*K + 1]
    lookup: np.ndarray  # SLOT_DTYPE
    coeffs: np.ndarray  # float64 [C_local]
    n_obs: int
    n_tallies: int
    data_bytes: int
    # PUB de-duplication: when some PUBs are the very same buffers as earlier ones,
    # ``k1_pubs`` lists the distinct PUBs -- the only ones staged and tallied -- while ``pubs`` keeps one row per
    # PUB for the product/sum kernel, duplicates carrying the offsets of their first occurrence.  None: no duplicates.
    k1_pubs: np.ndarray | None = None

    @property
    def staged_pubs(self) -> np.ndarray:
        """The PUBs whose bytes are staged and tallied (all of them unless duplicates were found)."""
        return self.pubs if self.k1_pubs is None else self.k1_pubs

    @property
    def dedup(self) -> dict:
        """How much the duplicate detection saved: PUBs / shot bytes handed over vs staged and tallied."""
        def nbytes(pubs):
            nb = self.groups["nb_obs"][pubs["group"]].astype(np.int64) + pubs["nb_qpd"].astype(np.int64)
            return int((pubs["shots"].astype(np.int64) * nb).sum())

        total, staged = nbytes(self.pubs), nbytes(self.staged_pubs)
        return {"pubs": int(len(self.pubs)), "distinct_pubs": int(len(self.staged_pubs)), "shot_bytes": total,
                "staged_shot_bytes": staged, "dedup_ratio": total / staged if staged else 1.0}

    @property
    def n_tiles(self) -> int:
        return int(((self.staged_pubs["shots"].astype(np.int64) + TILE_SHOTS - 1) // TILE_SHOTS).sum())

    @property
    def single_slot(self) -> bool:
        """Every observable has exactly one (group, term) slot in every subsystem (the usual case)."""
        return bool(np.all(np.diff(self.lookup_start.astype(np.int64)) == 1))

    @property
    def work(self) -> int:
        """shots x terms of this table (the unit of BASELINE.json's metric)."""
        return int((self.pubs["shots"].astype(np.int64) * self.groups["n_terms"][self.pubs["group"]]).sum())

    @property
    def algorithmic_bytes(self) -> int:
        """Bytes K1 + K2 must move: shot rows once, tallies written and read, coeffs, out."""
        pubs = self.staged_pubs
        nb = self.groups["nb_obs"][pubs["group"]].astype(np.int64) + pubs["nb_qpd"].astype(np.int64)
        shot_bytes = int((pubs["shots"].astype(np.int64) * nb).sum())
        return shot_bytes + 16 * self.n_tallies + 8 * len(self.coeffs) + 8 * self.n_obs


class TableBuilder:
    """Accumulates group variants, masks and lookups label by label."""

    def __init__(self, n_obs: int):
        self.n_obs = n_obs
        self.group_rows: list[tuple[int, int, int, int]] = []
        self.mask_chunks: list[np.ndarray] = []
        self.mask_bytes = 0
        self.variant: dict[tuple[int, int, int], int] = {}
        self.label_rows: list[tuple[int, int, int]] = []
        self.lookup_start: list[int] = [0]
        self.lookup_rows: list[tuple[int, int]] = []
        self._cogs: list[list] = []

    def add_label(self, collection, pub_base: int) -> int:
        """Register a subsystem; returns its index."""
        li = len(self.label_rows)
        self.label_rows.append((pub_base, len(collection.groups), li * self.n_obs))
        for slots in collection.lookup_indices:
            self.lookup_rows.extend(slots)
            self.lookup_start.append(len(self.lookup_rows))
        self._cogs.append(collection.groups)
        return li

    def group_variant(self, li: int, g: int, nb_obs: int) -> int:
        key = (li, g, nb_obs)
        v = self.variant.get(key)
        if v is None:
            cog = self._cogs[li][g]
            m = cog.mask_bytes(nb_obs)
            v = self.variant[key] = len(self.group_rows)
            self.group_rows.append((m.shape[0], nb_obs, self.mask_bytes, 0))
            self.mask_chunks.append(m.reshape(-1))
            self.mask_bytes += m.size
        return v

    def n_terms(self, li: int, g: int) -> int:
        return len(self._cogs[li][g].pauli_bitmasks)

    def tables(self, pubs: np.ndarray, coeffs: np.ndarray, n_tallies: int, data_bytes: int) -> HostTables:
        t = HostTables(
            groups=np.array(self.group_rows, dtype=GROUP_DTYPE),
            masks=np.concatenate(self.mask_chunks) if self.mask_chunks else np.zeros(0, dtype=np.uint8),
            pubs=pubs,
            labels=np.array(self.label_rows, dtype=LABEL_DTYPE),
            lookup_start=np.array(self.lookup_start, dtype=np.uint32),
            lookup=np.array(self.lookup_rows, dtype=SLOT_DTYPE).reshape(-1),
            coeffs=np.ascontiguousarray(coeffs, dtype=np.float64),
            n_obs=self.n_obs,
            n_tallies=n_tallies,
            data_bytes=data_bytes,
        )
        return t


# ------------------------------------------------------------------------------------------------
# Device batch: tables + shot bytes resident in HBM, and the two launches
# ------------------------------------------------------------------------------------------------
class Schedule:
    """PUB table in HBM + per-kernel tile lists (``qcut_schedule``)."""

    def __init__(self, plan: Plan, pubs: np.ndarray):
        self.pubs = np.ascontiguousarray(pubs, dtype=PUB_DTYPE)
        handle = C.c_void_p()
        check(
            _lib.load().qcut_schedule_create(
                plan.handle, _lib.np_ptr(self.pubs) if len(self.pubs) else None, len(self.pubs), C.byref(handle)
            ),
            "qcut_schedule_create",
        )
        self.handle = handle
        self.n_tiles = _lib.load().qcut_schedule_num_tiles(handle)
        self.d_pubs = _lib.load().qcut_schedule_device_pubs(handle)

    def __del__(self):
        handle, self.handle = getattr(self, "handle", None), None
        if handle:
            try:
                _lib.load().qcut_schedule_destroy(handle)
            except Exception:  # interpreter shutdown
                pass


class DeviceBatch:
    """One rank's share of a reconstruction, resident in HBM."""

    def __init__(self, tables: HostTables, data, *, flags: int = 0, k1_pubs: np.ndarray | None = None,
                 rounding: str = "exact", public_call: bool = False):
        """``k1_pubs`` (tally-table sharding): the PUB pieces THIS rank tallies -- descriptors with this rank's
        data offsets and shot counts but the global ``tally_offset`` -- while ``tables.pubs`` stays the complete
        table the product/sum kernel reads after the tally table has been all-reduced.

        ``rounding``: ``"exact"`` -- int64 tallies, ``E = tally / shots`` by one correctly rounded division, tuples
        summed in the reference's order up to 1024 of them (a fixed chunking above); ``"reference"`` -- every
        (PUB, term) chain accumulated shot by shot in fp64 and the tuples summed strictly in order, i.e. the
        reference's own arithmetic (``cutting_reconstruction.py:159,168``), bit for bit."""
        if rounding not in ("exact", "reference"):
            raise ValueError("rounding must be 'exact' or 'reference'")
        self.rounding = rounding
        self.tables = tables
        if k1_pubs is None:
            k1_pubs = tables.k1_pubs
        if public_call and flags == 0 and tables.data_bytes < _SMALL_CALL_BYTES and not _SYNC_JIT:
            flags = _lib.PLAN_NO_JIT  # launch-bound call: ahead-of-time kernels, no NVRTC at all
        self.plan = get_plan(tables.groups, tables.masks, flags=flags, background=public_call)
        if not public_call:
            self.plan.wait()  # resident / benchmark use: always the specialised kernels
        self.schedule = Schedule(self.plan, tables.pubs if k1_pubs is None else k1_pubs)
        self.data = data
        small = [tables.labels, tables.lookup_start, tables.lookup, tables.coeffs]
        if k1_pubs is not None:
            small.append(tables.pubs)
        dev = _upload_packed(small)
        self.labels, self.lookup_start, self.lookup, self.coeffs = dev[:4]
        self.k2_pubs = dev[4] if k1_pubs is not None else None
        # one spare element each: the error flag of the collectives (distributed.all_reduce_with_flag)
        self.tallies = torch.empty(tables.n_tallies + 1, dtype=torch.int64, device="cuda")
        self.out = torch.empty(tables.n_obs + 1, dtype=torch.float64, device="cuda")
        ws = _lib.load().qcut_reduce_workspace_bytes(len(tables.coeffs), tables.n_obs)
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device="cuda")
        self.tally_launches = 0
        self.expvals = None  # fp64 per-(PUB, term) expectation values of the reference-rounding mode

    @classmethod
    def from_skeleton(cls, skel: "_Skeleton", data: torch.Tensor, coeffs_dev: torch.Tensor, tables: HostTables,
                      rounding: str) -> "DeviceBatch":
        """A batch that shares everything layout-dependent (plan, schedule, uploaded tables) with earlier calls of the
        same shape; only the shot bytes, the coefficients (staged behind the shot bytes) and the outputs are its own."""
        self = cls.__new__(cls)
        self.rounding, self.tables, self.plan, self.schedule, self.data = rounding, tables, skel.plan, skel.schedule, data
        self.labels, self.lookup_start, self.lookup, self.coeffs = skel.labels, skel.lookup_start, skel.lookup, coeffs_dev
        self.k2_pubs = None
        self.tallies = torch.empty(tables.n_tallies + 1, dtype=torch.int64, device="cuda")
        self.out = torch.empty(tables.n_obs + 1, dtype=torch.float64, device="cuda")
        self.workspace = torch.empty(skel.workspace_bytes, dtype=torch.uint8, device="cuda")
        self.tally_launches = 0
        self.expvals = None
        self._single_slot = skel.single_slot
        return self

    def launch_tally(self) -> int:
        """K1 on the current stream: exact int64 sign tallies of every (PUB, term).  Returns #kernels launched."""
        n = C.c_int(0)
        check(
            _lib.load().qcut_tally_launch(
                self.plan.handle,
                self.schedule.handle,
                self.data.data_ptr(),
                self.tallies.data_ptr(),
                self.tables.n_tallies,
                _stream_ptr(),
                C.byref(n),
            ),
            "qcut_tally_launch",
        )
        self.tally_launches = n.value
        return n.value

    def launch_literal(self) -> int:
        """K1 with the reference's rounding on the current stream: fp64 ``expvals[pub.tally_offset + t]``."""
        if self.expvals is None:
            self.expvals = torch.empty(self.tables.n_tallies + 1, dtype=torch.float64, device="cuda")
        check(
            _lib.load().qcut_expvals_literal_launch(
                self.plan.handle, self.schedule.handle, self.data.data_ptr(), self.expvals.data_ptr(),
                self.tables.n_tallies, _stream_ptr(),
            ),
            "qcut_expvals_literal_launch",
        )
        return 1

    def launch_reduce(self) -> int:
        """K2 on the current stream: this rank's partial ``out[k]``.  Returns #kernels launched (2)."""
        t = self.tables
        literal = self.rounding == "reference"
        single_slot = getattr(self, "_single_slot", None)
        if single_slot is None:
            single_slot = self._single_slot = t.single_slot
        flags = (_lib.REDUCE_SINGLE_SLOT if single_slot else 0) | (_lib.REDUCE_SEQUENTIAL if literal else 0)
        fn = _lib.load().qcut_reduce_expvals_launch if literal else _lib.load().qcut_reduce_launch
        check(
            fn(
                self.expvals.data_ptr() if literal else self.tallies.data_ptr(),
                self.schedule.d_pubs if self.k2_pubs is None else self.k2_pubs.data_ptr(),
                self.labels.data_ptr(),
                len(t.labels),
                self.lookup_start.data_ptr(),
                self.lookup.data_ptr(),
                self.coeffs.data_ptr(),
                len(t.coeffs),
                t.n_obs,
                self.out.data_ptr(),
                self.workspace.data_ptr(),
                flags,
                _stream_ptr(),
            ),
            "qcut_reduce_launch",
        )
        if t.n_obs <= 0:
            return 0
        n_chunks = 1 if (literal or len(t.coeffs) <= 1024) else -(-len(t.coeffs) // 256)
        return 3 if n_chunks > 32 else 2

    def run(self) -> torch.Tensor:
        if self.rounding == "reference":
            self.launch_literal()
        else:
            self.launch_tally()
        self.launch_reduce()
        return self.out[: self.tables.n_obs]

    def run_tally_sharded(self, process_group, world: int) -> torch.Tensor:
        """Tally-table sharding: K1 on this rank's pieces, all-reduce of the int64 table (+ error flag), K2 on the
        complete table.  Every rank ends up with the full, bit-identical result."""
        from . import distributed

        self.launch_tally()
        self.tallies[-1] = 0
        distributed.all_reduce_with_flag(self.tallies, False, process_group, world)
        self.launch_reduce()
        self.out[-1] = self.tallies[-1].to(torch.float64)
        return self.out


# ------------------------------------------------------------------------------------------------
# Ingestion of SamplerV2 results: borrowed BitArray buffers -> pinned staging -> HBM
# ------------------------------------------------------------------------------------------------
_pinned: dict[int, torch.Tensor] = {}
_staging_locks: dict[int, threading.Lock] = {}
_staging_locks_guard = threading.Lock()


def _staging_lock(device: int) -> threading.Lock:
    with _staging_locks_guard:
        return _staging_locks.setdefault(device, threading.Lock())


def _pinned_buffer(nbytes: int) -> torch.Tensor:
    dev = torch.cuda.current_device()
    buf = _pinned.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _pinned.pop(dev, None)
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _pinned[dev] = buf
    return buf


def _collect_label(result, lo: int, hi: int):
    """Arrays of PUBs ``lo .. hi-1`` of one subsystem: (keepalive list, obs ptr, qpd ptr, shots, nb_obs, nb_qpd).

    The attribute walk ``result[idx].data.<register>.array`` and the address / shape of every array are
    read through the C API (``csrc/pyhelper.c``); arrays that are not C-contiguous ``uint8`` matrices, or
    whose row counts differ, take the per-PUB Python path below.
    """
    n = hi - lo
    keep: list = []
    obs_ptr = np.zeros(n, dtype=np.uint64)
    qpd_ptr = np.zeros(n, dtype=np.uint64)
    shots = np.zeros(n, dtype=np.int64)
    nb_obs = np.zeros(n, dtype=np.int64)
    nb_qpd = np.zeros(n, dtype=np.int64)
    flags = np.zeros(n, dtype=np.uint8)
    if n:
        got = _lib.load_pyhelper().qcut_py_collect(
            result, lo, hi, obs_ptr.ctypes.data, qpd_ptr.ctypes.data, shots.ctypes.data, nb_obs.ctypes.data,
            nb_qpd.ctypes.data, flags.ctypes.data, keep,
        )
        if got != n:
            raise RuntimeError("qcut_py_collect failed")
    u8 = np.dtype(np.uint8)
    for j in np.flatnonzero(flags):
        o, q = np.asarray(keep[2 * j]), np.asarray(keep[2 * j + 1])
        if o.ndim != 2 or q.ndim != 2:
            raise ValueError("Only PUBs of shape () are supported: BitArray.array must be (num_shots, num_bytes).")
        s = q.shape[0]
        if o.shape[0] < s:
            # the reference indexes obs_array[j] for j < qpd_array.shape[0] (cutting_reconstruction.py:155-157)
            raise IndexError(f"index {o.shape[0]} is out of bounds for axis 0 with size {o.shape[0]}")
        o = np.ascontiguousarray(o[:s], dtype=u8)
        q = np.ascontiguousarray(q, dtype=u8)
        keep[2 * j], keep[2 * j + 1] = o, q
        obs_ptr[j], qpd_ptr[j] = o.ctypes.data, q.ctypes.data
        shots[j], nb_obs[j], nb_qpd[j] = s, o.shape[1], q.shape[1]
    return keep, obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd


@dataclass
class HostArrays:
    """Borrowed shot buffers of a call, as flat address tables (obs_0, qpd_0, obs_1, qpd_1, ...)."""

    keep: list  # the array objects: keeps the addresses valid
    src: np.ndarray  # uint64 host addresses
    nbytes: np.ndarray  # uint64 sizes
    dst: np.ndarray  # uint64 offsets in the staged buffer

    def __len__(self) -> int:
        return len(self.src)


_DEDUP = os.environ.get("QCUT_DEDUP", "1") not in ("0", "", "off")


class _DuplicateFinder:
    """Finds PUBs whose two ``BitArray`` buffers are the very same memory as an earlier PUB's (``qcut_dedup_*``).

    The reference emits one subexperiment per (sample, group) and label even when the label's own map ids repeat
    (``cutting_experiments.py:146-157``; ``docs/explanation/index.rst:188``), so a caller who runs each distinct
    circuit once hands the same result object over many times.  Same addresses + same shape + same group variant =>
    same bytes, same masks => same tallies: such a PUB is neither staged nor tallied again.  Content is never
    hashed (two separately sampled runs of one circuit differ)."""

    _idle: list = []  # finders of finished calls: their hash tables keep their (already touched) memory

    @classmethod
    def acquire(cls) -> "_DuplicateFinder":
        try:
            finder = cls._idle.pop()
        except IndexError:
            return cls()
        _lib.load().qcut_dedup_clear(finder.handle)
        finder.duplicates = 0
        return finder

    def release(self) -> None:
        if len(self._idle) < 4:
            self._idle.append(self)

    def __init__(self):
        handle = C.c_void_p()
        check(_lib.load().qcut_dedup_create(C.byref(handle)), "qcut_dedup_create")
        self.handle = handle
        self.duplicates = 0

    def representatives(self, base: int, obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd, variant) -> np.ndarray:
        """Global index of the first PUB with the same buffers, for every PUB of a block starting at ``base``."""
        n = len(obs_ptr)
        rep = np.empty(n, dtype=np.int64)
        cols = [np.ascontiguousarray(a, dtype=t) for a, t in ((obs_ptr, np.uint64), (qpd_ptr, np.uint64), (shots, np.int64),
                                                              (nb_obs, np.int64), (nb_qpd, np.int64), (variant, np.uint32))]
        found = _lib.load().qcut_dedup_block(self.handle, base, *[_lib.np_ptr(a) for a in cols], n, _lib.np_ptr(rep))
        if found < 0:
            check(1, "qcut_dedup_block")
        self.duplicates += int(found)
        return rep

    def __del__(self):
        handle, self.handle = getattr(self, "handle", None), None
        if handle:
            try:
                _lib.load().qcut_dedup_destroy(handle)
            except Exception:  # interpreter shutdown
                pass


def build_host_tables(results_by_label: list, collections: list, coeffs: np.ndarray, i_lo: int, i_hi: int,
                      n_obs: int, *, sink=None, block_pubs: int = 0, collected: dict | None = None) -> tuple[HostTables, HostArrays]:
    """Walk the PUBs of coefficient range ``[i_lo, i_hi)`` of every label (host only, no device).

    Offsets, tally slots and group variants are computed with NumPy, not per PUB.  The walk proceeds in
    blocks of ``block_pubs`` PUBs (0: one block per label; with a sink the blocks of a label grow 4x each).  Without ``sink`` the blocks are laid out back to
    back in one staging buffer.  With ``sink`` every block is handed over as soon as it has been walked
    (``sink(src, nbytes, dst, block_bytes)`` returns the block's byte offset from a base address the sink
    fixes later through ``sink.rebase()``): staging of block b overlaps the walk of block b+1.
    """
    builder = TableBuilder(n_obs)
    keep_all: list = []
    src_chunks, nbytes_chunks, dst_chunks, chunks, bases = [], [], [], [], []
    finder = _DuplicateFinder.acquire() if _DEDUP else None
    k1_chunks, rep_chunks = [], []  # distinct PUBs of every block / representative (global PUB index) of every PUB
    offset = 0
    tally_offset = 0
    n_pubs = 0
    for result, collection in zip(results_by_label, collections):
        n_groups = len(collection.groups)
        li = builder.add_label(collection, n_pubs)
        walked = collected.get(li) if collected else None  # (keep, obs_ptr, ...) of ALL PUBs of the label
        terms_of_group = np.array([builder.n_terms(li, g) for g in range(n_groups)], dtype=np.int64)
        p_lo, p_hi = i_lo * n_groups, i_hi * n_groups
        step = block_pubs if block_pubs > 0 else max(p_hi - p_lo, 1)
        b_hi = p_lo
        while b_hi < p_hi:
            b_lo, b_hi = b_hi, min(b_hi + step, p_hi)
            if sink is not None:
                step *= 4  # every hand-over costs the sink a pipeline fill: few, growing blocks
            if walked is None:
                keep, obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd = _collect_label(result, b_lo, b_hi)
            else:
                keep = walked[0][2 * b_lo: 2 * b_hi]
                obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd = (a[b_lo:b_hi] for a in walked[1:])
            n = b_hi - b_lo
            g_idx = np.arange(b_lo, b_hi, dtype=np.int64) % n_groups
            # group variants: one per distinct (group, row width)
            variant = np.empty(n, dtype=np.uint32)
            key = (g_idx << 32) + nb_obs
            for k in np.unique(key).tolist():
                variant[key == k] = builder.group_variant(li, k >> 32, k & 0xFFFFFFFF)
            terms = terms_of_group[g_idx]
            if n and int(shots.max()) >= 2**31:
                raise ValueError("A PUB with 2**31 or more shots is not supported.")
            pubs = np.zeros(n, dtype=PUB_DTYPE)
            pubs["shots"] = shots
            pubs["group"] = variant
            pubs["nb_qpd"] = nb_qpd
            # PUBs that are the same buffers as an earlier PUB are neither staged nor tallied again
            rep = finder.representatives(n_pubs, obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd, variant) if finder else None
            rep_chunks.append(rep if rep is not None else n_pubs + np.arange(n, dtype=np.int64))
            fresh = None
            if rep is not None and n:
                fresh = rep == n_pubs + np.arange(n, dtype=np.int64)
                if fresh.all():
                    fresh = None
            if fresh is not None:
                obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd, terms = (a[fresh] for a in (obs_ptr, qpd_ptr, shots, nb_obs, nb_qpd, terms))
            m = len(shots)  # distinct PUBs of the block
            raw = np.empty(2 * m, dtype=np.int64)
            raw[0::2] = shots * nb_obs
            raw[1::2] = shots * nb_qpd
            sizes = (raw + _ALIGN - 1) // _ALIGN * _ALIGN
            starts = np.cumsum(sizes) - sizes
            block_bytes = int(sizes.sum())
            src = np.empty(2 * m, dtype=np.uint64)
            src[0::2] = obs_ptr
            src[1::2] = qpd_ptr
            nbytes = raw.astype(np.uint64)
            if sink is None:
                base = offset
            else:
                base = 0
                bases.append(sink(src, nbytes, starts.astype(np.uint64), block_bytes))
            k1 = pubs if fresh is None else pubs[fresh]
            k1["obs_offset"] = base + starts[0::2]
            k1["qpd_offset"] = base + starts[1::2]
            k1["tally_offset"] = tally_offset + np.cumsum(terms) - terms
            src_chunks.append(src)
            nbytes_chunks.append(nbytes)
            dst_chunks.append((base + starts).astype(np.uint64))
            offset += block_bytes
            tally_offset += int(terms.sum())
            n_pubs += n
            chunks.append(pubs)
            k1_chunks.append(k1)
            keep_all.extend(keep)
    if sink is not None:
        # the blocks live in separate device allocations: make every offset relative to the lowest one
        for k1, delta in zip(k1_chunks, sink.rebase(bases)):
            k1["obs_offset"] += np.uint64(delta)
            k1["qpd_offset"] += np.uint64(delta)
    pubs = np.concatenate(chunks) if chunks else np.zeros(0, dtype=PUB_DTYPE)
    k1_pubs = None
    if finder is not None and finder.duplicates:
        # one row per PUB for the product/sum kernel: a duplicate carries the offsets of its first occurrence
        k1_pubs = np.concatenate(k1_chunks)
        rep_all = np.concatenate(rep_chunks)
        fresh_all = rep_all == np.arange(len(pubs))
        for f in ("obs_offset", "qpd_offset", "tally_offset"):
            pubs[f][fresh_all] = k1_pubs[f]
            pubs[f] = pubs[f][rep_all]
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint64)  # noqa: E731
    arrays = HostArrays(keep_all, cat(src_chunks), cat(nbytes_chunks), cat(dst_chunks))
    if finder is not None:
        finder.release()
    tables = builder.tables(pubs, coeffs[i_lo:i_hi], tally_offset, max(offset, _ALIGN))
    tables.k1_pubs = k1_pubs
    return tables, arrays


def gather_arrays(tables: HostTables, arrays: HostArrays, dst_ptr: int, n_threads: int = 0) -> None:
    """Copy every borrowed buffer to its offset in the staging buffer at ``dst_ptr`` (threads in C)."""
    if len(arrays) == 0:
        return
    check(
        _lib.load().qcut_gather_host(
            _lib.np_ptr(arrays.src), _lib.np_ptr(arrays.nbytes), _lib.np_ptr(arrays.dst), len(arrays), dst_ptr, n_threads
        ),
        "qcut_gather_host",
    )


# Pinned staging ring of the end-to-end path, independent of the problem size.  Measured on cfg4 (14.3 GB per
# call, PCIe ceiling of the box 55 GB/s).  One rank per host: 292 ms with 256 MiB written with streaming stores
# (qcut_stage_h2d does that for rings >= 64 MiB per call), 308 ms with 128 MiB / 326 ms with 64 or 256 MiB and
# ordinary stores, 352 ms with 32 MiB (pieces too small for PCIe).  Several ranks on one host: the path is bound by
# the host's memory-copy ceiling -- every shot byte is read from its BitArray and written to the ring (then read by
# the DMA engine).  Round 2 sweep with two ranks on a 24-core host whose copy ceiling (scripts/host_bw.py, 16
# threads) is 60 GB/s: 64 MiB ring x 8 threads 30.3 GB/s per rank (= the ceiling), 16 MiB x 4 30.5, 8 MiB x 2 22.5,
# 4 MiB x 2 20.1, 2 MiB x 1 11.4 -- a ring small enough to stay in the caches does NOT help (the copy itself is
# the limit, not the DMA read), so the defaults stay: 256 MiB alone on a host, 64 MiB with local peers.
def _local_peers() -> int:
    try:
        return max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        return 1


def _default_staging_bytes() -> int:
    return (256 if _local_peers() <= 1 else 64) << 20


_STAGING_BYTES = int(os.environ.get("QCUT_STAGING_MB", "0") or 0) << 20 or _default_staging_bytes()
# Gather threads per feeder (two feeders).  4 and 8 measured the same alone on a host; with two ranks on a 24-core
# host 8 (64 MiB ring) and 4 (16 MiB) were the best of a sweep, 2 and 1 clearly worse.
_STAGING_THREADS = int(os.environ.get("QCUT_STAGING_THREADS", "0") or 0) or 8
_BLOCK_PUBS = 2048  # PUBs in the first staging block of a label (~1 ms of attribute walk); later blocks grow 4x


class _BlockStager:
    """Background staging of walked blocks: one thread feeds ``qcut_stage_h2d`` while the caller keeps walking.

    Every block gets its own device allocation (its size is only known once it has been walked); ``rebase``
    turns the block addresses into offsets from the lowest one, which becomes the kernels' ``d_data``.
    The ctypes call releases the GIL, so the gather + PCIe transfer of block b run beside the C-API walk of
    block b+1.  Copies are queued on the stream that was current when the stager was created.
    """

    def __init__(self):
        self.device = torch.cuda.current_device()
        self.stream = _stream_ptr()
        self.pinned = _pinned_buffer(_STAGING_BYTES)
        self.blocks: list[torch.Tensor] = []
        self.base = 0
        self.error: BaseException | None = None
        self.block_seconds: list[tuple[int, float, float]] = []  # (bytes, start, end) of every staged block
        self.queue: queue.SimpleQueue = queue.SimpleQueue()
        # two feeders, each with half of the ring: the pipeline fill of one block (gather before the first
        # copy) overlaps the drain of the previous one
        half = (self.pinned.numel() // 2) & ~4095
        self.threads = [
            threading.Thread(target=self._run, args=(self.pinned.data_ptr() + i * half, half), name=f"qcut-stage-{i}",
                             daemon=True)
            for i in range(2)
        ]
        for t in self.threads:
            t.start()

    def __call__(self, src: np.ndarray, nbytes: np.ndarray, dst: np.ndarray, block_bytes: int) -> int:
        block = torch.empty(block_bytes + _lib.DATA_SLACK, dtype=torch.uint8, device="cuda")
        self.blocks.append(block)
        if len(src):
            self.queue.put((src, nbytes, dst, block.data_ptr()))
        return block.data_ptr()

    def rebase(self, addresses: list[int]) -> list[int]:
        self.base = min(addresses) if addresses else 0
        return [a - self.base for a in addresses]

    def _run(self, ring_ptr: int, ring_bytes: int) -> None:
        try:
            torch.cuda.set_device(self.device)  # qcut_stage_h2d copies to the calling thread's device
            lib = _lib.load()
            while True:
                item = self.queue.get()
                if item is None:
                    return
                if self.error is not None:
                    continue  # drain
                src, nbytes, dst, d_ptr = item
                t0 = time.perf_counter()
                check(
                    lib.qcut_stage_h2d(_lib.np_ptr(src), _lib.np_ptr(nbytes), _lib.np_ptr(dst), len(src),
                                       ring_ptr, ring_bytes, d_ptr, _STAGING_THREADS, self.stream),
                    "qcut_stage_h2d",
                )
                self.block_seconds.append((int(nbytes.sum()), t0, time.perf_counter()))
        except BaseException as ex:  # noqa: BLE001 - re-raised on the caller's thread by finish()
            self.error = ex

    def finish(self, *, raise_errors: bool = True) -> None:
        """Wait until every queued block is in HBM (``qcut_stage_h2d`` returns after its last copy)."""
        for _ in self.threads:
            self.queue.put(None)
        for t in self.threads:
            t.join()
        if raise_errors and self.error is not None:
            raise self.error

    def data_ptr(self) -> int:
        return self.base


_SIMPLE_PUBS = 8192          # up to this many PUBs the whole call is walked first ...
_SIMPLE_BYTES = 64 << 20      # ... and staged by one gather + one H2D copy if it is at most this large
_copy_done: dict[int, torch.cuda.Event] = {}  # per device: last H2D copy out of the pinned buffer


_PIPELINE_FROM = 4 << 20      # simple path: from this size on the gather and the H2D copy are pipelined in 2 MiB pieces


def _stage_simple(tables: HostTables, arrays: HostArrays, extra: np.ndarray | None = None) -> torch.Tensor:
    """Small inputs: gather every borrowed buffer into the pinned buffer, ONE host-to-device copy on the current
    stream (no threads, nothing pipelined: for tutorial-sized calls the pipeline's fixed costs were the larger part
    of the call).  From 4 MiB on, ``qcut_stage_h2d`` gathers 2 MiB pieces on eight threads and copies each as soon as
    it is complete (cfg2, 30 MB: 1.0 ms with four threads).  ``extra`` (the coefficients of a cached layout) travels behind the shot bytes, at
    ``_align(data_bytes + DATA_SLACK)``, in the same copy."""
    dev = torch.cuda.current_device()
    n = tables.data_bytes
    extra_off = _align(n + _lib.DATA_SLACK)
    total = extra_off + (extra.nbytes if extra is not None else 0)
    pinned = _pinned_buffer(max(total, 16 << 20))
    done = _copy_done.get(dev)
    if done is not None:
        done.synchronize()  # a previous call's copy may still be reading the pinned buffer
    data = torch.empty(max(total, n + _lib.DATA_SLACK), dtype=torch.uint8, device="cuda")
    if n >= _PIPELINE_FROM and len(arrays):
        check(
            _lib.load().qcut_stage_h2d(_lib.np_ptr(arrays.src), _lib.np_ptr(arrays.nbytes), _lib.np_ptr(arrays.dst), len(arrays),
                                       pinned.data_ptr(), 16 << 20, data.data_ptr(), 8, _stream_ptr()),
            "qcut_stage_h2d",
        )  # returns when the bytes have left the pinned pieces
        if extra is not None:
            data[extra_off: total].copy_(torch.from_numpy(extra.view(np.uint8).reshape(-1)), non_blocking=False)
        return data
    gather_arrays(tables, arrays, pinned.data_ptr(), n_threads=1)
    if extra is not None:
        np.frombuffer((C.c_uint8 * extra.nbytes).from_address(pinned.data_ptr() + extra_off), dtype=np.uint8)[:] = \
            extra.view(np.uint8).reshape(-1)
        data[:total].copy_(pinned[:total], non_blocking=True)
    else:
        data[:n].copy_(pinned[:n], non_blocking=True)
    ev = _copy_done.get(dev) or torch.cuda.Event()
    ev.record()
    _copy_done[dev] = ev
    return data


@dataclass
class _Skeleton:
    """Everything about a small call that depends only on its LAYOUT (observable groups, PUB shapes): repeated calls
    of one shape -- the tutorial loop, a sweep over circuits of one family -- skip table building, the schedule upload
    and four small copies, and stage their coefficients behind the shot bytes."""

    tables: HostTables
    nbytes: np.ndarray
    dst: np.ndarray
    plan: Plan
    schedule: Schedule
    labels: torch.Tensor
    lookup_start: torch.Tensor
    lookup: torch.Tensor
    workspace_bytes: int
    single_slot: bool
    collections: list  # keeps the ids in the signature unique
    dedup: dict = field(default_factory=dict)  # tables.dedup of the layout (no duplicates by construction)


_skeletons: dict = {}
_skeleton_lock = threading.Lock()
_SKELETON_CACHE = os.environ.get("QCUT_LAYOUT_CACHE", "1") not in ("0", "", "off")


def _stage_small(results_by_label: list, collections: list, coeffs: np.ndarray, i_lo: int, i_hi: int, n_obs: int,
                 flags: int, rounding: str, public_call: bool, t_start: float):
    """The simple path with the layout cache.  Returns a DeviceBatch, or None when the call is not small after all
    (more than ``_SIMPLE_BYTES``), in which case ``walked`` has been consumed for nothing but a few microseconds."""
    global last_stage_trace
    dev = torch.cuda.current_device()
    walked = {li: _collect_label(result, i_lo * len(c.groups), i_hi * len(c.groups))
              for li, (result, c) in enumerate(zip(results_by_label, collections))}
    sig = None
    if _SKELETON_CACHE:
        sig = (dev, flags, public_call, n_obs, i_hi - i_lo, tuple(id(c) for c in collections),
               tuple(a.tobytes() for w in walked.values() for a in w[3:]))
        with _skeleton_lock:
            skel = _skeletons.get(sig)
        if skel is not None:
            finder = _DuplicateFinder.acquire() if _DEDUP else None
            dup, base = 0, 0
            if finder is not None:
                for li, w in walked.items():
                    variant = np.zeros(len(w[1]), dtype=np.uint32) + np.uint32(li)  # any per-label constant will do here
                    finder.representatives(base, w[1], w[2], w[3], w[4], w[5], variant)
                    base += len(w[1])
                dup = finder.duplicates
                finder.release()
            if dup == 0:
                src = np.empty(len(skel.nbytes), dtype=np.uint64)
                keep, at = [], 0
                for w in walked.values():
                    m = len(w[1])
                    src[at: at + 2 * m: 2] = w[1]
                    src[at + 1: at + 2 * m: 2] = w[2]
                    at += 2 * m
                    keep.extend(w[0])
                arrays = HostArrays(keep, src, skel.nbytes, skel.dst)
                local = np.ascontiguousarray(coeffs[i_lo:i_hi], dtype=np.float64)
                t_walked = time.perf_counter()
                with _staging_lock(dev):
                    data = _stage_simple(skel.tables, arrays, extra=local)
                off = _align(skel.tables.data_bytes + _lib.DATA_SLACK)
                import dataclasses

                tables = dataclasses.replace(skel.tables, coeffs=local)
                batch = DeviceBatch.from_skeleton(skel, data, data[off: off + local.nbytes], tables, rounding)
                batch.keepalive = keep
                last_stage_trace = {"walk_s": t_walked - t_start, "stage_s": time.perf_counter() - t_walked, "tables_s": 0.0,
                                    "path": "simple (cached layout)", "dedup": skel.dedup}
                return batch
    tables, arrays = build_host_tables(results_by_label, collections, coeffs, i_lo, i_hi, n_obs, collected=walked)
    if tables.data_bytes > _SIMPLE_BYTES:
        return None
    t_walked = time.perf_counter()
    with _staging_lock(dev):
        data = _stage_simple(tables, arrays)
    t_staged = time.perf_counter()
    batch = DeviceBatch(tables, data, flags=flags, rounding=rounding, public_call=public_call)
    batch.keepalive = arrays.keep
    last_stage_trace = {"walk_s": t_walked - t_start, "stage_s": t_staged - t_walked,
                        "tables_s": time.perf_counter() - t_staged, "path": "simple", "dedup": tables.dedup}
    if sig is not None and tables.k1_pubs is None and batch.k2_pubs is None:
        skel = _Skeleton(tables, arrays.nbytes, arrays.dst, batch.plan, batch.schedule, batch.labels, batch.lookup_start,
                         batch.lookup, max(_lib.load().qcut_reduce_workspace_bytes(len(tables.coeffs), tables.n_obs), 16),
                         tables.single_slot, list(collections), tables.dedup)
        with _skeleton_lock:
            if len(_skeletons) >= 32:
                _skeletons.pop(next(iter(_skeletons)))
            _skeletons[sig] = skel
    return batch


def stage_v2(results_by_label: list, collections: list, coeffs: np.ndarray, i_lo: int, i_hi: int, n_obs: int,
             *, flags: int = 0, collected: dict | None = None, rounding: str = "exact",
             public_call: bool = False) -> DeviceBatch:
    """Stage the PUBs of coefficient range ``[i_lo, i_hi)`` of every label into HBM.

    Large inputs: the PUBs are walked block by block (``csrc/pyhelper.c``); background threads hand every walked
    block to ``qcut_stage_h2d``, whose worker threads gather runs of borrowed ``BitArray`` buffers into a small
    pinned ring and queue one H2D copy per run: attribute walk, host gather and PCIe transfer all overlap.
    Small inputs (<= 8192 PUBs and <= 64 MiB): one walk, one gather, one copy (``_stage_simple``).
    """
    _lib.require_device()
    t_start = time.perf_counter()
    n_pubs = (i_hi - i_lo) * sum(len(c.groups) for c in collections)
    global last_stage_trace
    if n_pubs <= _SIMPLE_PUBS and collected is None:
        batch = _stage_small(results_by_label, collections, coeffs, i_lo, i_hi, n_obs, flags, rounding, public_call, t_start)
        if batch is not None:
            return batch
    # The pinned ring is shared by the calls of a process on one device: concurrent callers (Python threads)
    # take turns for the staging phase -- it is PCIe bound anyway -- and run their kernels independently.
    with _staging_lock(torch.cuda.current_device()):
        stager = _BlockStager()
        try:
            tables, _ = build_host_tables(results_by_label, collections, coeffs, i_lo, i_hi, n_obs, sink=stager,
                                          block_pubs=_BLOCK_PUBS, collected=collected)
        except BaseException:
            stager.finish(raise_errors=False)
            raise
        # plan lookup / schedule upload proceed while the last blocks are still in flight
        t_walked = time.perf_counter()
        try:
            batch = DeviceBatch(tables, stager, flags=flags, rounding=rounding, public_call=public_call)
        finally:
            t_batch = time.perf_counter()
            stager.finish(raise_errors=False)
    if stager.error is not None:
        raise stager.error
    last_stage_trace = {
        "walk_s": t_walked - t_start, "tables_s": t_batch - t_walked, "wait_s": time.perf_counter() - t_batch,
        "blocks": [(b, s - t_start, e - t_start) for b, s, e in stager.block_seconds], "path": "pipelined",
        "dedup": tables.dedup,
    }
    return batch


def stage_v2_pieces(results_by_label: list, collections: list, coeffs: np.ndarray, n_obs: int, rank: int,
                    world: int, *, flags: int = 0, collected: dict | None = None, public_call: bool = False) -> DeviceBatch:
    """Tally-table sharding (``distributed`` module docstring): every rank describes ALL PUBs (descriptors only),
    stages the bytes of its own byte-balanced piece of the flattened (label, PUB, shot range) list and tallies it
    into the global tally table."""
    from . import distributed

    _lib.require_device()
    tables, arrays = build_host_tables(results_by_label, collections, coeffs, 0, len(coeffs), n_obs, collected=collected)
    pubs = tables.staged_pubs  # duplicates are neither staged nor tallied
    nb_obs = tables.groups["nb_obs"][pubs["group"]].astype(np.int64) if len(pubs) else np.zeros(0, dtype=np.int64)
    nb_qpd = pubs["nb_qpd"].astype(np.int64)
    pieces = distributed.piece_ranges(nb_obs + nb_qpd, pubs["shots"], world, TILE_SHOTS)[rank]
    k1 = np.zeros(len(pieces), dtype=PUB_DTYPE)
    src = np.zeros(2 * len(pieces), dtype=np.uint64)
    nbytes = np.zeros(2 * len(pieces), dtype=np.uint64)
    dst = np.zeros(2 * len(pieces), dtype=np.uint64)
    offset = 0
    for j, (p, lo, hi) in enumerate(pieces):
        n = hi - lo
        o_bytes, q_bytes = n * int(nb_obs[p]), n * int(nb_qpd[p])
        k1[j] = pubs[p]
        k1[j]["shots"] = n
        k1[j]["obs_offset"] = offset
        src[2 * j], nbytes[2 * j], dst[2 * j] = int(arrays.src[2 * p]) + lo * int(nb_obs[p]), o_bytes, offset
        offset += _align(o_bytes)
        k1[j]["qpd_offset"] = offset
        src[2 * j + 1], nbytes[2 * j + 1], dst[2 * j + 1] = int(arrays.src[2 * p + 1]) + lo * int(nb_qpd[p]), q_bytes, offset
        offset += _align(q_bytes)
    piece_arrays = HostArrays(arrays.keep, src, nbytes, dst)
    local = HostTables(groups=tables.groups, masks=tables.masks, pubs=k1, labels=tables.labels,
                       lookup_start=tables.lookup_start, lookup=tables.lookup, coeffs=tables.coeffs, n_obs=n_obs,
                       n_tallies=tables.n_tallies, data_bytes=max(offset, _ALIGN))
    with _staging_lock(torch.cuda.current_device()):
        if local.data_bytes <= _SIMPLE_BYTES:
            data = _stage_simple(local, piece_arrays)
        else:
            stager = _BlockStager()
            try:
                stager.rebase([stager(src, nbytes, dst, local.data_bytes)])
            finally:
                stager.finish(raise_errors=False)
            if stager.error is not None:
                raise stager.error
            data = stager
    batch = DeviceBatch(tables, data, flags=flags, k1_pubs=k1, public_call=public_call)
    batch.keepalive = arrays.keep
    return batch


def _label_shapes(result, n: int):
    """Walk PUBs 0..n-1 of one label; returns the tuple ``_collect_label`` returns."""
    return _collect_label(result, 0, n)


def plan_sharding(results_by_label: list, collections: list, n_coeffs: int, rank: int, world: int,
                  *, allow_tally_mode: bool = True):
    """Decide how a call is cut over ``world`` ranks.  Deterministic: every rank reaches the same decision from the
    same inputs (integer arithmetic on shapes every rank can see)."""
    from . import distributed as D

    if world <= 1:
        return D.Sharding("coefficients", 0, 1, 0, n_coeffs)
    lo, hi = D.coefficient_slice(n_coeffs, rank, world)
    plain = D.Sharding("coefficients", rank, world, lo, hi)
    if any(isinstance(r, D.ShardedResult) or hasattr(r, "quasi_dists") for r in results_by_label) or n_coeffs == 0:
        # every rank holds its own equal-count shard / SamplerV1 quasi-distributions (no shot bytes): nothing to
        # balance, nothing to look at
        return plain
    groups = [len(c.groups) for c in collections]
    n_pubs = n_coeffs * sum(groups)
    if n_pubs > D.FULL_WALK_PUBS:
        # look at a sample of every label first: equal shapes -> equal-count slices are byte balanced
        uniform = True
        for result, G in zip(results_by_label, groups):
            shapes = set()
            for g in range(G):
                for i in np.unique(np.linspace(0, n_coeffs - 1, num=min(n_coeffs, 32), dtype=np.int64)).tolist():
                    _, _, _, s, nbo, nbq = _collect_label(result, i * G + g, i * G + g + 1)
                    shapes.add((g, int(s[0]), int(nbo[0]), int(nbq[0])))
            uniform = uniform and len(shapes) == G
        if uniform:
            return plain
    collected = {}
    weights = np.zeros(n_coeffs, dtype=np.int64)
    for li, (result, G) in enumerate(zip(results_by_label, groups)):
        walked = _label_shapes(result, n_coeffs * G)
        collected[li] = walked
        _, _, _, shots, nbo, nbq = walked
        weights += (shots * (nbo + nbq)).reshape(n_coeffs, G).sum(axis=1)
    slices = D.balanced_slices(weights, world)
    imb = D.imbalance(weights, slices)
    n_tallies = n_coeffs * sum(len(g.pauli_bitmasks) for c in collections for g in c.groups)
    if allow_tally_mode and (n_coeffs < world or imb > D.MAX_IMBALANCE) and n_tallies <= D.MAX_TALLY_TABLE:
        return D.Sharding("tallies", rank, world, 0, n_coeffs, "bytes", 1.0, collected)
    lo, hi = slices[rank]
    return D.Sharding("coefficients", rank, world, lo, hi, "bytes", imb, collected)


last_stage_trace: dict = {}  # timing of the most recent stage_v2 call (diagnostics only)
