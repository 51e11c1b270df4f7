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
import threading
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
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    if raw.size == 0:
        return torch.zeros(16, dtype=torch.uint8, device="cuda")
    return torch.from_numpy(raw.copy()).to("cuda", non_blocking=False)


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


def get_plan(groups: np.ndarray, masks: np.ndarray, *, flags: int = 0) -> Plan:
    key = (torch.cuda.current_device(), flags, groups.tobytes(), masks.tobytes())
    with _plan_lock:
        plan = _plan_cache.get(key)
        if plan is None:
            if len(_plan_cache) >= 16:
                _plan_cache.pop(next(iter(_plan_cache)))
            plan = _plan_cache[key] = Plan(groups, masks, flags=flags)
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

    @property
    def n_tiles(self) -> int:
        return int(((self.pubs["shots"].astype(np.int64) + TILE_SHOTS - 1) // TILE_SHOTS).sum())

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
        nb = self.groups["nb_obs"][self.pubs["group"]].astype(np.int64) + self.pubs["nb_qpd"].astype(np.int64)
        shot_bytes = int((self.pubs["shots"].astype(np.int64) * nb).sum())
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

    def __init__(self, tables: HostTables, data: torch.Tensor, *, flags: int = 0):
        self.tables = tables
        self.plan = get_plan(tables.groups, tables.masks, flags=flags)
        self.schedule = Schedule(self.plan, tables.pubs)
        self.data = data
        self.labels = _to_device(tables.labels)
        self.lookup_start = _to_device(tables.lookup_start)
        self.lookup = _to_device(tables.lookup)
        self.coeffs = _to_device(tables.coeffs)
        self.tallies = torch.empty(max(tables.n_tallies, 1), dtype=torch.int64, device="cuda")
        self.out = torch.empty(max(tables.n_obs, 1), dtype=torch.float64, device="cuda")
        ws = _lib.load().qcut_reduce_workspace_bytes(len(tables.coeffs), tables.n_obs)
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device="cuda")
        self.tally_launches = 0

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

    def launch_reduce(self) -> int:
        """K2 on the current stream: this rank's partial ``out[k]``.  Returns #kernels launched (2)."""
        t = self.tables
        check(
            _lib.load().qcut_reduce_launch(
                self.tallies.data_ptr(),
                self.schedule.d_pubs,
                self.labels.data_ptr(),
                len(t.labels),
                self.lookup_start.data_ptr(),
                self.lookup.data_ptr(),
                self.coeffs.data_ptr(),
                len(t.coeffs),
                t.n_obs,
                self.out.data_ptr(),
                self.workspace.data_ptr(),
                int(t.single_slot),
                _stream_ptr(),
            ),
            "qcut_reduce_launch",
        )
        return 2 if t.n_obs > 0 else 0

    def run(self) -> torch.Tensor:
        self.launch_tally()
        self.launch_reduce()
        return self.out[: self.tables.n_obs]


# ------------------------------------------------------------------------------------------------
# Ingestion of SamplerV2 results: borrowed BitArray buffers -> pinned staging -> HBM
# ------------------------------------------------------------------------------------------------
_pinned: dict[int, torch.Tensor] = {}


def _pinned_buffer(nbytes: int) -> torch.Tensor:
    dev = torch.cuda.current_device()
    buf = _pinned.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _pinned.pop(dev, None)
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _pinned[dev] = buf
    return buf


def _pub_arrays(pub) -> tuple[np.ndarray, np.ndarray]:
    data = pub.data
    obs = data.observable_measurements.array
    qpd = data.qpd_measurements.array
    if obs.ndim != 2 or qpd.ndim != 2:
        raise ValueError("Only PUBs of shape () are supported: BitArray.array must be (num_shots, num_bytes).")
    shots = qpd.shape[0]
    if obs.shape[0] < shots:
        # the reference indexes obs_array[j] for j < qpd_array.shape[0] (cutting_reconstruction.py:155-157)
        raise IndexError(
            f"index {obs.shape[0]} is out of bounds for axis 0 with size {obs.shape[0]}"
        )
    if obs.shape[0] != shots:
        obs = obs[:shots]
    if obs.dtype != np.uint8 or not obs.flags.c_contiguous:
        obs = np.ascontiguousarray(obs, dtype=np.uint8)
    if qpd.dtype != np.uint8 or not qpd.flags.c_contiguous:
        qpd = np.ascontiguousarray(qpd, dtype=np.uint8)
    return obs, qpd


def build_host_tables(results_by_label: list, collections: list, coeffs: np.ndarray, i_lo: int, i_hi: int,
                      n_obs: int) -> tuple[HostTables, list[np.ndarray]]:
    """Walk the PUBs of coefficient range ``[i_lo, i_hi)`` of every label (host only, no device).

    Returns the tables and the borrowed ``[obs_0, qpd_0, obs_1, qpd_1, ...]`` arrays in PUB order.
    """
    builder = TableBuilder(n_obs)
    arrays: list[np.ndarray] = []
    pub_rows = []
    offset = 0
    tally_offset = 0
    for result, collection in zip(results_by_label, collections):
        n_groups = len(collection.groups)
        li = builder.add_label(collection, len(pub_rows))
        for i in range(i_lo, i_hi):
            for g in range(n_groups):
                obs, qpd = _pub_arrays(result[i * n_groups + g])
                shots, nb_obs = obs.shape
                nb_qpd = qpd.shape[1]
                obs_off = offset
                offset += _align(shots * nb_obs)
                qpd_off = offset
                offset += _align(shots * nb_qpd)
                arrays.append(obs)
                arrays.append(qpd)
                pub_rows.append(
                    (obs_off, qpd_off, tally_offset, shots, builder.group_variant(li, g, nb_obs), nb_qpd, 0)
                )
                tally_offset += builder.n_terms(li, g)
    pubs = np.array(pub_rows, dtype=PUB_DTYPE) if pub_rows else np.zeros(0, dtype=PUB_DTYPE)
    if len(pubs) and int(pubs["shots"].max()) >= 2**31:
        raise ValueError("A PUB with 2**31 or more shots is not supported.")
    return builder.tables(pubs, coeffs[i_lo:i_hi], tally_offset, max(offset, _ALIGN)), arrays


def gather_arrays(tables: HostTables, arrays: list[np.ndarray], dst_ptr: int, n_threads: int = 0) -> None:
    """Copy every borrowed buffer to its offset in the staging buffer at ``dst_ptr`` (threads in C)."""
    n = len(arrays)
    if n == 0:
        return
    src = np.fromiter((a.ctypes.data for a in arrays), dtype=np.uint64, count=n)
    nbytes = np.fromiter((a.nbytes for a in arrays), dtype=np.uint64, count=n)
    dst = np.empty(n, dtype=np.uint64)
    dst[0::2] = tables.pubs["obs_offset"]
    dst[1::2] = tables.pubs["qpd_offset"]
    check(
        _lib.load().qcut_gather_host(_lib.np_ptr(src), _lib.np_ptr(nbytes), _lib.np_ptr(dst), n, dst_ptr, n_threads),
        "qcut_gather_host",
    )


def stage_v2(results_by_label: list, collections: list, coeffs: np.ndarray, i_lo: int, i_hi: int, n_obs: int,
             *, flags: int = 0) -> DeviceBatch:
    """Stage the PUBs of coefficient range ``[i_lo, i_hi)`` of every label into HBM."""
    _lib.require_device()
    tables, arrays = build_host_tables(results_by_label, collections, coeffs, i_lo, i_hi, n_obs)
    # Gather every borrowed buffer into one pinned staging buffer (threads), then one H2D copy.
    total = tables.data_bytes
    pinned = _pinned_buffer(total)
    gather_arrays(tables, arrays, pinned.data_ptr())
    data = torch.empty(total, dtype=torch.uint8, device="cuda")
    data.copy_(pinned[:total], non_blocking=True)
    return DeviceBatch(tables, data, flags=flags)
