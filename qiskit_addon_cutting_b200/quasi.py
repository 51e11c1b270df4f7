"""SamplerV1 (quasi-distribution) results on the device, alone or mixed with SamplerV2 subsystems.

Reference: ``cutting_reconstruction.py:143-149`` with ``_process_outcome`` (``:173-193``) and
``_outcome_to_int`` (``:225-232``).  Each outcome key is split on the host into its observable
bits (low ``max(1, len(pauli_indices))`` bits) and qpd bits; the kernel accumulates
``weight * sign`` per (PUB, term) sequentially in dict order, which reproduces the reference's
fp64 result bit for bit; K2 then runs on those expectation values.

The reference decides V1 / V2 per subsystem (``:143,150``), so one call may mix them.  The SamplerV2
subsystems then go through the normal ingest and K1 (or the reference-rounding chain kernel), their int64 tallies
are turned into fp64 expectation values on the device (``qcut_tallies_to_expvals_launch``: one correctly rounded
division, what K2 would have done), and ONE K2 launch consumes the merged table -- subsystems in the caller's
order, because the product over subsystems is formed in that order.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import GROUP_DTYPE, LABEL_DTYPE, PUB_DTYPE, SLOT_DTYPE, check
from .engine import _stream_ptr, _to_device


def _outcome_to_int(outcome) -> int:
    if isinstance(outcome, (int, np.integer)):
        return int(outcome)
    outcome = outcome.replace(" ", "")
    if len(outcome) < 2 or outcome[1] in ("0", "1"):
        return int(f"0b{outcome}", 0)
    return int(outcome, 0)


def _limbs(value: int, n: int) -> list[int]:
    return [(value >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(n)]


def _v1_expvals(results: list, collections: list, i_lo: int, i_hi: int):
    """Per-(PUB, term) expectation values of the SamplerV1 subsystems: (pubs table, pub_base per subsystem, fp64
    device vector).  ``pubs[p].tally_offset`` indexes the vector; PUBs are listed subsystem after subsystem, ``i * G + g``."""
    lib = _lib.load()
    pubs_raw = []  # (global group, [(obs, qpd, weight)])
    group_rows, bases = [], []
    mask_ints: list[list[int]] = []
    max_obs_bits, max_qpd_bits = 1, 1
    for result, coll in zip(results, collections):
        n_groups = len(coll.groups)
        bases.append(len(pubs_raw))
        g_base = len(mask_ints)
        for cog in coll.groups:
            mask_ints.append(list(cog.pauli_bitmasks))
            max_obs_bits = max(max_obs_bits, cog.num_measured_bits)
        for i in range(i_lo, i_hi):
            for g, cog in enumerate(coll.groups):
                nmb = cog.num_measured_bits
                items = []
                for outcome, weight in result.quasi_dists[i * n_groups + g].items():
                    o = _outcome_to_int(outcome)
                    obs, qpd = o & ((1 << nmb) - 1), o >> nmb
                    max_qpd_bits = max(max_qpd_bits, qpd.bit_length())
                    items.append((obs, qpd, float(weight)))
                pubs_raw.append((g_base + g, items))
    n_lo, n_lq = (max_obs_bits + 63) // 64, (max_qpd_bits + 63) // 64

    mask_limbs, offset = [], 0
    for masks in mask_ints:
        group_rows.append((len(masks), 0, offset, 0))
        for m in masks:
            mask_limbs.extend(_limbs(m, n_lo))
        offset += len(masks) * n_lo
    pubs = np.zeros(len(pubs_raw), dtype=PUB_DTYPE)
    dist_start = np.zeros(len(pubs_raw) + 1, dtype=np.int64)
    obs_limbs, qpd_limbs, weights = [], [], []
    tally_offset = 0
    for p, (group, items) in enumerate(pubs_raw):
        pubs[p]["group"] = group
        pubs[p]["tally_offset"] = tally_offset
        tally_offset += len(mask_ints[group])
        for obs, qpd, w in items:
            obs_limbs.extend(_limbs(obs, n_lo))
            qpd_limbs.extend(_limbs(qpd, n_lq))
            weights.append(w)
        dist_start[p + 1] = len(weights)

    expvals = torch.zeros(max(tally_offset, 1), dtype=torch.float64, device="cuda")
    if len(pubs):
        d_pubs = _to_device(pubs)
        d_groups = _to_device(np.array(group_rows, dtype=GROUP_DTYPE))
        d_masks = _to_device(np.array(mask_limbs, dtype=np.uint64))
        d_dist = _to_device(dist_start)
        d_obs = _to_device(np.array(obs_limbs, dtype=np.uint64))
        d_qpd = _to_device(np.array(qpd_limbs, dtype=np.uint64))
        d_w = _to_device(np.array(weights, dtype=np.float64))
        check(
            lib.qcut_quasi_launch(
                d_pubs.data_ptr(), len(pubs), d_groups.data_ptr(), d_masks.data_ptr(), d_dist.data_ptr(),
                d_obs.data_ptr(), n_lo, d_qpd.data_ptr(), n_lq, d_w.data_ptr(), expvals.data_ptr(), _stream_ptr(),
            ),
            "qcut_quasi_launch",
        )
    return pubs, bases, expvals[:tally_offset]


def reconstruct_v1(results: list, collections: list, coeffs: np.ndarray, i_lo: int, i_hi: int, n_obs: int,
                   *, sequential: bool = False, v1: list | None = None, rounding: str = "exact") -> torch.Tensor:
    """Partial ``out[k]`` of coefficient range ``[i_lo, i_hi)`` for SamplerV1 inputs, or for a mix of SamplerV1 and
    SamplerV2 subsystems (``v1[l]`` says which; default: all V1)."""
    from . import engine

    _lib.require_device()
    lib = _lib.load()
    if v1 is None:
        v1 = [True] * len(results)
    idx1 = [li for li, f in enumerate(v1) if f]
    idx2 = [li for li, f in enumerate(v1) if not f]
    sequential = sequential or rounding == "reference"

    # SamplerV2 subsystems: normal ingest, K1 (or the reference-rounding chains) -> fp64 per (PUB, term)
    pubs2, bases2, e2, batch2 = np.zeros(0, dtype=PUB_DTYPE), [], None, None
    if idx2:
        batch2 = engine.stage_v2([results[li] for li in idx2], [collections[li] for li in idx2], coeffs, i_lo, i_hi, n_obs,
                                 rounding=rounding, public_call=True)
        t2 = batch2.tables
        if rounding == "reference":
            batch2.launch_literal()
            e2 = batch2.expvals[: t2.n_tallies]
        else:
            batch2.launch_tally()
            e2 = torch.empty(max(t2.n_tallies, 1), dtype=torch.float64, device="cuda")
            check(
                lib.qcut_tallies_to_expvals_launch(
                    batch2.plan.handle, batch2.schedule.d_pubs if batch2.k2_pubs is None else batch2.k2_pubs.data_ptr(),
                    len(t2.pubs), batch2.tallies.data_ptr(), e2.data_ptr(), _stream_ptr()),
                "qcut_tallies_to_expvals_launch",
            )
            e2 = e2[: t2.n_tallies]
        pubs2 = t2.pubs
        bases2 = [int(lab["pub_base"]) for lab in t2.labels]
    n2 = 0 if e2 is None else int(e2.numel())

    pubs1, bases1, e1 = _v1_expvals([results[li] for li in idx1], [collections[li] for li in idx1], i_lo, i_hi)
    pubs1 = pubs1.copy()
    pubs1["tally_offset"] += np.uint64(n2)

    # merged tables, subsystems in the caller's order (the product over subsystems is formed in that order, :163-166)
    label_rows, lookup_start, lookup_rows = [], [0], []
    for li, coll in enumerate(collections):
        if v1[li]:
            base = len(pubs2) + bases1[idx1.index(li)]
        else:
            base = bases2[idx2.index(li)]
        label_rows.append((base, len(coll.groups), li * n_obs))
        for slots in coll.lookup_indices:
            lookup_rows.extend(slots)
            lookup_start.append(len(lookup_rows))
    pubs = np.concatenate([pubs2, pubs1]) if len(pubs2) else pubs1
    expvals = e1 if e2 is None else torch.cat([e2, e1])
    if expvals.numel() == 0:
        expvals = torch.zeros(1, dtype=torch.float64, device="cuda")

    d_pubs = _to_device(pubs)
    d_labels = _to_device(np.array(label_rows, dtype=LABEL_DTYPE))
    d_ls = _to_device(np.array(lookup_start, dtype=np.uint32))
    d_lk = _to_device(np.array(lookup_rows, dtype=SLOT_DTYPE).reshape(-1))
    d_coeffs = _to_device(np.ascontiguousarray(coeffs[i_lo:i_hi], dtype=np.float64))
    out = torch.empty(max(n_obs, 1), dtype=torch.float64, device="cuda")
    ws = torch.empty(max(lib.qcut_reduce_workspace_bytes(i_hi - i_lo, n_obs), 16), dtype=torch.uint8, device="cuda")
    check(
        lib.qcut_reduce_expvals_launch(
            expvals.data_ptr(), d_pubs.data_ptr(), d_labels.data_ptr(), len(label_rows), d_ls.data_ptr(),
            d_lk.data_ptr(), d_coeffs.data_ptr(), i_hi - i_lo, n_obs, out.data_ptr(), ws.data_ptr(),
            (_lib.REDUCE_SINGLE_SLOT if np.all(np.diff(np.array(lookup_start, dtype=np.int64)) == 1) else 0)
            | (_lib.REDUCE_SEQUENTIAL if sequential else 0), _stream_ptr(),
        ),
        "qcut_reduce_expvals_launch",
    )
    return out[:n_obs]
