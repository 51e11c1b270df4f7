"""Parity at BASELINE-sized workloads (B200): data generated on the device, checked through properties that do
not need the slow oracle on every byte:

* the three K1 implementations (NVRTC-specialised, run-time-mask bit-sliced, generic AND/POPC) are independent
  code paths and must produce identical int64 tallies on the same resident bytes;
* a seeded subsample of PUBs is copied back and recomputed by the C oracle (bit-exact);
* identity: a term with an empty mask tallies S - 2 * #(shots with odd qpd parity);
* the full expectation values of two runs (two schedules, same data) are bit-identical (determinism).
"""

import numpy as np
import pytest
import torch

import workloads as W
from oracle import c_oracle
from qiskit_addon_cutting_b200 import _lib, engine

pytestmark = pytest.mark.gpu


def _run(tables, data, flags):
    batch = engine.DeviceBatch(tables, data, flags=flags)
    out = batch.run().clone()
    torch.cuda.synchronize()
    return batch, batch.tallies[: tables.n_tallies].clone(), out


@pytest.mark.parametrize("name,scale", [("cfg4_heavy_hex", 0.1), ("cfg5_molecular", 0.15), ("cfg3_wide_dense", 0.3),
                                        ("cfg3_wide_local", 1.0), ("cfg2_wire", 1.0), ("cfg1_tutorial", 1.0),
                                        ("cfg4u_heavy_hex_unpartitioned", 0.02)])
def test_three_kernels_agree_and_match_oracle_subsample(lib, name, scale):
    w = W.get_workload(name, scale)
    tables = W.build_tables(w)
    data = W.device_data(tables, seed=w.seed)
    b_jit, t_jit, out_jit = _run(tables, data, 0)
    b_rt, t_rt, out_rt = _run(tables, data, _lib.PLAN_NO_JIT)
    assert torch.equal(t_jit, t_rt), "specialised vs run-time-mask kernel"
    assert max(b_jit.plan.kernels()) >= _lib.KERNEL_JIT0 and max(b_rt.plan.kernels()) <= _lib.KERNEL_SLICED_RT
    if tables.work <= 3e10:  # the generic kernel is POPC bound: keep it to a few hundred ms
        _, t_gen, out_gen = _run(tables, data, _lib.PLAN_GENERIC_ONLY)
        assert torch.equal(t_jit, t_gen), "specialised vs generic kernel"
        assert torch.equal(out_jit, out_gen)
    assert torch.equal(out_jit, out_rt)
    # determinism across runs
    _, t_again, out_again = _run(tables, data, 0)
    assert torch.equal(t_jit, t_again) and torch.equal(out_jit, out_again)

    # seeded subsample of PUBs vs the C oracle
    rng = np.random.default_rng(w.seed)
    picks = rng.choice(len(tables.pubs), size=min(24, len(tables.pubs)), replace=False)
    t_host = t_jit.cpu().numpy()
    for p in picks:
        pub = tables.pubs[p]
        grp = tables.groups[pub["group"]]
        nb, T, S, nbq = int(grp["nb_obs"]), int(grp["n_terms"]), int(pub["shots"]), int(pub["nb_qpd"])
        S = min(S, 200_000)  # bound the CPU work for 10^6-shot PUBs: compare a prefix via a separate tally below
        obs = data[int(pub["obs_offset"]): int(pub["obs_offset"]) + S * nb].cpu().numpy().reshape(S, nb)
        qpd = data[int(pub["qpd_offset"]): int(pub["qpd_offset"]) + S * nbq].cpu().numpy().reshape(S, nbq)
        masks = tables.masks[int(grp["mask_offset"]): int(grp["mask_offset"]) + T * nb].reshape(T, nb)
        _, want = c_oracle.pub(obs, qpd, masks)
        if S == int(pub["shots"]):
            assert np.array_equal(t_host[int(pub["tally_offset"]): int(pub["tally_offset"]) + T], want), (name, int(p))
        # identity / empty-mask terms: tally = S - 2 * #(odd qpd parity)
        empty = np.flatnonzero(~masks.any(axis=1))
        if empty.size and S == int(pub["shots"]):
            odd = int((np.bitwise_count(qpd).sum(axis=1) & 1).sum())
            assert t_host[int(pub["tally_offset"]) + int(empty[0])] == S - 2 * odd


def test_prefix_tallies_of_million_shot_pubs(lib):
    """cfg3 shape (10^6 shots per PUB): a PUB truncated to its first S' shots must tally like the oracle on that prefix."""
    w = W.get_workload("cfg3_wide_dense", 0.06)  # 2 tuples x 2 subsystems
    tables = W.build_tables(w)
    data = W.device_data(tables, seed=7)
    for shots in (4096 * 3 + 17, 150_001):
        t2 = W.build_tables(w)
        t2.pubs["shots"] = shots
        _, tallies, _ = _run(t2, data, 0)
        tallies = tallies.cpu().numpy()
        pub = t2.pubs[1]
        grp = t2.groups[pub["group"]]
        nb, T = int(grp["nb_obs"]), int(grp["n_terms"])
        obs = data[int(pub["obs_offset"]): int(pub["obs_offset"]) + shots * nb].cpu().numpy().reshape(shots, nb)
        qpd = data[int(pub["qpd_offset"]): int(pub["qpd_offset"]) + shots].cpu().numpy().reshape(shots, 1)
        masks = t2.masks[int(grp["mask_offset"]): int(grp["mask_offset"]) + T * nb].reshape(T, nb)
        _, want = c_oracle.pub(obs, qpd, masks)
        assert np.array_equal(tallies[int(pub["tally_offset"]): int(pub["tally_offset"]) + T], want)


@pytest.mark.parametrize("shots,exact", [(1 << 20, True), (1_000_000, False)])
def test_expectation_values_at_a_million_shots_vs_literal_oracle(lib, shots, exact):
    """cfg3 shot count through the whole path (K1 + K2) against the C restatement of the reference's arithmetic
    (sequential fp64 accumulation of +-fl(1/S) per shot, cutting_reconstruction.py:159).

    S = 2^20: 1/S and every partial sum are exact, so the reference has no rounding noise and our result must be
    bit-identical to it.  S = 10^6: the reference's own accumulation deviates from tally/S by up to ~6e-12 per
    expectation value; we require |ours - reference arithmetic| <= 2e-11 * kappa and
    |ours - exact| <= 4 ulp * kappa.
    """
    w = W.get_workload("cfg3_wide_dense", 0.06)  # 2 coefficient tuples x 2 subsystems x 100 terms
    w.shots = shots
    tables = W.build_tables(w)
    data = W.device_data(tables, seed=11)
    batch, tallies, out = _run(tables, data, 0)
    host = data[: tables.data_bytes].cpu().numpy()
    e_seq, tally = c_oracle.pubs(host, tables.pubs, tables.groups, tables.masks, tables.n_tallies, want_seq=True, n_threads=0)
    assert np.array_equal(tallies.cpu().numpy(), tally)

    def reduce(E):  # reference order: product over subsystems, sum over tuples (cutting_reconstruction.py:134-168)
        res = np.zeros(tables.n_obs)
        for i, c in enumerate(tables.coeffs):
            cur = np.ones(tables.n_obs)
            for lab in tables.labels:
                for k in range(tables.n_obs):
                    e0 = int(tables.lookup_start[int(lab["lookup_base"]) + k])
                    slot = tables.lookup[e0]
                    pub = tables.pubs[int(lab["pub_base"]) + i * int(lab["num_groups"]) + int(slot["group"])]
                    cur[k] *= E[int(pub["tally_offset"]) + int(slot["term"])]
            res += c * cur
        return res

    ours = out.cpu().numpy()
    kappa = float(np.abs(tables.coeffs).sum())
    exact_vals = reduce(tally / float(shots))
    literal_vals = reduce(e_seq)
    assert np.max(np.abs(ours - exact_vals)) <= 4 * np.finfo(np.float64).eps * kappa
    assert [v.hex() for v in ours] == [v.hex() for v in exact_vals]
    if exact:
        assert [v.hex() for v in ours] == [v.hex() for v in literal_vals]
    else:
        assert np.max(np.abs(ours - literal_vals)) <= 2e-11 * kappa
    # rounding="reference": the chain kernel reproduces the reference's arithmetic bit for bit at 10^6 shots, both per
    # (PUB, term) and after the product / sum
    lit = engine.DeviceBatch(tables, data, rounding="reference")
    lit_out = lit.run().cpu().numpy()
    assert [v.hex() for v in lit.expvals[: tables.n_tallies].cpu().numpy()] == [v.hex() for v in e_seq]
    assert [v.hex() for v in lit_out] == [v.hex() for v in literal_vals]
    print(f"\nS={shots}: max|ours-literal|={np.max(np.abs(ours - literal_vals)):.3e} max|ours-exact|="
          f"{np.max(np.abs(ours - exact_vals)):.3e} max|literal-exact|={np.max(np.abs(literal_vals - exact_vals)):.3e} "
          f"max|E|={np.max(np.abs(exact_vals)):.3e} reference-rounding mode: bit-identical")


def _host_reduce_inputs(tables, values):
    """term[i, k] = c_i * prod_label E[label, i, k] from a per-(PUB, term) value table, vectorised (single-slot lookups),
    multiplied in the reference's order: ones * E_label0 * E_label1 * ... then * coeff (cutting_reconstruction.py:135,163-168)."""
    C, K = len(tables.coeffs), tables.n_obs
    cur = np.ones((C, K))
    i = np.arange(C, dtype=np.int64)[:, None]
    for lab in tables.labels:
        base, G, lb = int(lab["pub_base"]), int(lab["num_groups"]), int(lab["lookup_base"])
        slots = tables.lookup[tables.lookup_start[lb: lb + K]]
        pub = base + i * G + slots["group"].astype(np.int64)[None, :]
        idx = tables.pubs["tally_offset"].astype(np.int64)[pub] + slots["term"].astype(np.int64)[None, :]
        cur = cur * values[idx]
    return tables.coeffs[:, None] * cur


@pytest.mark.parametrize("name", ["cfg4_heavy_hex", "cfg5_molecular"])
def test_full_size_shapes_against_the_oracle(lib, name):
    """BASELINE-sized cfg4 / cfg5 (scale 1.0, biased data).  Tallies: the specialised kernels against the run-time-mask
    kernels on every PUB and against the C oracle on a sample.  Expectation values: the default mode against a NumPy
    reduction of the exact tallies; the reference-rounding mode BIT-IDENTICAL to the strictly sequential NumPy sum
    (cumsum) of its own per-PUB values, which in turn are bit-identical to the C restatement of the reference's
    shot-by-shot accumulation on the sample."""
    w = W.get_workload(name, 1.0)
    tables = W.build_tables(w)
    assert tables.single_slot
    data = W.device_data(tables, seed=w.seed)
    batch, tallies, out = _run(tables, data, 0)
    _, t_rt, out_rt = _run(tables, data, _lib.PLAN_NO_JIT)
    assert torch.equal(tallies, t_rt) and torch.equal(out, out_rt)
    t_host = tallies.cpu().numpy()
    shots = tables.pubs["shots"].astype(np.float64)
    T_of = tables.groups["n_terms"][tables.pubs["group"]].astype(np.int64)
    exact_e = t_host / np.repeat(shots, T_of)
    terms = _host_reduce_inputs(tables, exact_e)
    kappa = float(np.abs(tables.coeffs).sum())
    ours = out.cpu().numpy()
    assert np.max(np.abs(ours - terms.sum(axis=0))) <= 1e-11 * max(kappa, 1.0)
    # reference rounding at full size
    lit = engine.DeviceBatch(tables, data, rounding="reference")
    lit_out = lit.run().cpu().numpy()
    e_lit = lit.expvals[: tables.n_tallies].cpu().numpy()
    seq = np.cumsum(_host_reduce_inputs(tables, e_lit), axis=0)[-1]  # strictly sequential, like expvals += ... (:168)
    assert [v.hex() for v in lit_out] == [v.hex() for v in seq]
    # sample of PUBs against the C oracle: int64 tallies and the reference's shot-by-shot fp64 values
    rng = np.random.default_rng(w.seed + 1)
    worst = 0.0
    for p in rng.choice(len(tables.pubs), size=24, replace=False):
        pub = tables.pubs[p]
        grp = tables.groups[pub["group"]]
        nb, T, S, nbq = int(grp["nb_obs"]), int(grp["n_terms"]), int(pub["shots"]), int(pub["nb_qpd"])
        obs = data[int(pub["obs_offset"]): int(pub["obs_offset"]) + S * nb].cpu().numpy().reshape(S, nb)
        qpd = data[int(pub["qpd_offset"]): int(pub["qpd_offset"]) + S * nbq].cpu().numpy().reshape(S, nbq)
        masks = tables.masks[int(grp["mask_offset"]): int(grp["mask_offset"]) + T * nb].reshape(T, nb)
        e_seq, want = c_oracle.pub(obs, qpd, masks)
        sl = slice(int(pub["tally_offset"]), int(pub["tally_offset"]) + T)
        assert np.array_equal(t_host[sl], want), (name, int(p))
        assert [v.hex() for v in e_lit[sl]] == [v.hex() for v in e_seq], (name, int(p))
        worst = max(worst, float(np.max(np.abs(e_seq - want / S))))
    print(f"\n{name} at scale 1.0: max|default - numpy reduction| = {np.max(np.abs(ours - terms.sum(axis=0))):.3e} (kappa {kappa:.0f}); "
          f"max|reference mode - default| = {np.max(np.abs(lit_out - ours)):.3e}; reference's per-PUB noise on the sample {worst:.3e}")
