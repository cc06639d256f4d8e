"""GPU parity through the C-ABI for everything around the heuristic period: RL actions, masks, CatsLog
transforms, both reward types, heterogeneous parameters, replay tapes, ragged / tiny / scaled
economies, the host-buffer entry point, and the reference-shaped `Cats` / `CatsLog` adapters.

Oracle = in-repo CPU restatement of docs/MODEL.md ("parity unpinned" w.r.t. ABCredit.jl).
Bar: int32 state equal, float64 state, observations and rewards equal to the last bit.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import bits, compare_state, make_oracles

pytestmark = pytest.mark.gpu


def _env(B, **kw):
    from r_mabm_b200.engine import BatchedCats
    return BatchedCats(B, **kw)


def _flags(log_mode=False, price_change="agent_price", reward_type="profits"):
    return ((orc.FLAG_LOG if log_mode else 0) | (orc.FLAG_AVG_PRICE if price_change == "avg_price" else 0)
            | (orc.FLAG_REWARD_RMS if reward_type == "rms" else 0))


@pytest.mark.parametrize("log_mode,price_change,reward_type", [
    (False, "agent_price", "profits"), (False, "avg_price", "rms"), (True, "agent_price", "profits"),
    (True, "avg_price", "rms")])
def test_rl_actions_obs_reward_bit_exact(log_mode, price_change, reward_type):
    """cats.py:434-464 / 673-704 (actions), 466-481 / 655-671 (obs), 493-556 (reward)."""
    B, A, W, F, N = 5, 3, 400, 40, 8
    env = _env(B, W=W, F=F, N=N, n_agents=A, log_mode=log_mode, price_change=price_change,
               reward_type=reward_type, seed=17, bankruptcy_reward=-7.5)
    orcs = make_oracles(B, W, F, N, seed=17, bankruptcy_reward=-7.5)
    env.run(30)
    for o in orcs:
        for _ in range(30):
            o.step()
    rng = np.random.default_rng(5)
    fl = _flags(log_mode, price_change, reward_type)
    for t in range(12):
        lo, hi = (-0.3, 0.3) if log_mode else (0.7, 1.3)
        acts = rng.uniform(lo, hi, size=(B, A, 2))
        mask = (rng.random((B, A)) < 0.7).astype(np.uint8)
        obs, rew, term, _ = env.step(torch.from_numpy(acts).cuda(), torch.from_numpy(mask).cuda())
        torch.cuda.synchronize()
        g_obs, g_rew = obs.cpu().numpy(), rew.cpu().numpy()
        for b, o in enumerate(orcs):
            r_obs, r_rew = o.step(actions=acts[b], mask=mask[b], flags=fl)
            assert np.array_equal(bits(g_obs[b]), bits(r_obs)), (t, b, g_obs[b], r_obs)
            assert np.array_equal(bits(g_rew[b]), bits(r_rew)), (t, b, g_rew[b], r_rew)
        assert compare_state(env, orcs, f"t={t}") == []


def test_zero_action_zero_price_survives():
    """reference tests/test_env.py:157-178 steps with actions [0, 0]: De -> alpha, P -> 0.0."""
    B, A = 2, 3
    env = _env(B, W=300, F=30, N=6, n_agents=A, seed=2)
    orcs = make_oracles(B, 300, 30, 6, seed=2)
    env.run(20)
    for o in orcs:
        for _ in range(20):
            o.step()
    acts = torch.zeros((B, A, 2), dtype=torch.float64, device="cuda")
    for t in range(5):
        obs, rew, term, _ = env.step(acts)
        for b, o in enumerate(orcs):
            r_obs, r_rew = o.step(actions=np.zeros((A, 2)), flags=0)
            assert np.array_equal(bits(obs[b].cpu().numpy()), bits(r_obs))
            assert np.array_equal(bits(rew[b].cpu().numpy()), bits(r_rew))
        assert compare_state(env, orcs, f"t={t}") == []
    assert torch.isfinite(env.field("scal")).all()


def test_heterogeneous_parameters_per_economy():
    """params as [B][34]: every economy follows its own row, z values differ."""
    B = 4
    plist = [dict(orc.DEFAULT_PARAMS) for _ in range(B)]
    plist[1].update(z_c=3, z_e=2, chi=0.08)
    plist[2].update(z_c=8, z_k=4, z_e=7, Iprob=0.5, subsidy=0.3, tax_rate=0.1, tax_rate_d=0.05)
    plist[3].update(z_c=1, z_k=1, z_e=1, maastricht=2.0, wage_update_up=0.2)
    env = _env(B, W=500, F=50, N=10, params=[dict(p) for p in plist], seed=23)
    orcs = make_oracles(B, 500, 50, 10, params=plist, seed=23)
    for t in range(4):
        env.run(10)
        orc.run_batch(orcs, 10, n_threads=4, want_stats=False)
        torch.cuda.synchronize()
        assert compare_state(env, orcs, f"t={10 * (t + 1)}") == []


def test_tape_replay_matches_oracle_and_philox():
    """Replay mode (north_star): draws injected from a tape give the same bits as the Philox source."""
    B, W, F, N, T = 3, 300, 30, 8, 12
    a, b = _env(B, W=W, F=F, N=N, seed=31), _env(B, W=W, F=F, N=N, seed=31)
    orcs = make_oracles(B, W, F, N, seed=31)
    for t in range(T):
        recs = np.stack([o.fill_tape() for o in orcs])           # [B, tape_bytes]
        assert recs.shape[1] == a.tape_bytes()
        tape = torch.from_numpy(recs).cuda().contiguous().view(-1)
        a.run(1, tape=tape)
        b.run(1)
        for o, r in zip(orcs, recs):
            o.step(tape=r)
        torch.cuda.synchronize()
        assert torch.equal(a.blobs, b.blobs), t
        assert compare_state(a, orcs, f"t={t}") == []
    # a multi-period tape in one launch
    c = _env(B, W=W, F=F, N=N, seed=31)
    d = make_oracles(B, W, F, N, seed=31)
    recs = np.zeros((B, 5, c.tape_bytes()), dtype=np.uint8)
    for t in range(5):
        for i, o in enumerate(d):
            recs[i, t] = o.fill_tape()
            o.step(tape=recs[i, t])
    c.run(5, tape=torch.from_numpy(recs).cuda().contiguous().view(-1))
    torch.cuda.synchronize()
    assert compare_state(c, d, "5-period tape") == []


@pytest.mark.parametrize("W,F,N,zs", [(37, 3, 1, (5, 2, 5)), (1, 1, 1, (5, 2, 5)), (64, 33, 31, (8, 8, 8)),
                                       (130, 7, 2, (2, 1, 3)), (96, 32, 32, (4, 4, 4))])
def test_ragged_and_tiny_economies(W, F, N, zs):
    """Sizes that are not multiples of 32, fewer firms than candidates (z clamps to n), single agents."""
    B = 3
    p = dict(orc.DEFAULT_PARAMS, z_c=zs[0], z_k=zs[1], z_e=zs[2])
    env = _env(B, W=W, F=F, N=N, params=dict(p), seed=41)
    orcs = make_oracles(B, W, F, N, params=p, seed=41)
    for t in range(3):
        env.run(15)
        orc.run_batch(orcs, 15, n_threads=3, want_stats=False)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"{(W, F, N)} t={15 * (t + 1)}")
        assert bad == [], "\n".join(bad)


def test_scaled_economy_10x():
    """BASELINE.json config 5: 10x households and firms (W=10000, F=1000, N=200)."""
    B, W, F, N = 2, 10000, 1000, 200
    env = _env(B, W=W, F=F, N=N, seed=51)
    orcs = make_oracles(B, W, F, N, seed=51)
    for t in range(2):
        env.run(8)
        orc.run_batch(orcs, 8, n_threads=2, want_stats=False)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"10x t={8 * (t + 1)}")
        assert bad == [], "\n".join(bad)


def test_step_host_equals_device_step():
    """cats_step_host (pinned host buffers, the `e2e` path of bench.py) == cats_step on device tensors."""
    B, A = 6, 2
    a, b = _env(B, n_agents=A, seed=61), _env(B, n_agents=A, seed=61)
    a.run(10); b.run(10)
    acts = torch.full((B, A, 2), 1.05, dtype=torch.float64)
    mask = torch.ones((B, A), dtype=torch.uint8)
    for _ in range(3):
        ho, hr, ht, hs = a.step_host(acts.pin_memory(), mask.pin_memory(), want_stats=True)
        do, dr, dt, ds = b.step(acts.cuda(), mask.cuda(), want_stats=True)
        torch.cuda.synchronize()
        assert torch.equal(ho, do.cpu()) and torch.equal(hr, dr.cpu()) and torch.equal(ht, dt.cpu())
        assert torch.equal(hs, ds.cpu())
    assert torch.equal(a.blobs, b.blobs)
    # the caller's own buffers (not laid out like the staging block): one copy per buffer, same results;
    # and the block's own input views, filled in place
    import ctypes as C
    from r_mabm_b200 import _native as nat
    h_obs = torch.empty((B, A, 2), dtype=torch.float64).pin_memory()
    h_rew = torch.empty((B, A), dtype=torch.float64).pin_memory()
    h_term = torch.empty((B,), dtype=torch.uint8).pin_memory()
    h_stats = torch.empty((B, 1, nat.N_STATS), dtype=torch.float64).pin_memory()
    args = a._args(1, a.flags, want_io=False)
    pa, pm = acts.pin_memory(), mask.pin_memory()
    args.d_actions, args.d_action_mask = pa.data_ptr(), pm.data_ptr()
    args.d_obs, args.d_reward, args.d_terminated = h_obs.data_ptr(), h_rew.data_ptr(), h_term.data_ptr()
    args.d_status, args.d_stats = None, h_stats.data_ptr()
    nat.check(a._lib.cats_step_host(C.byref(args), a._host_scratch.data_ptr()))
    ia, im = b.host_action_buffers()
    ia.copy_(acts); im.copy_(mask)
    ho, hr, ht, hs = b.step_host(ia, im, want_stats=True)
    assert torch.equal(h_obs, ho) and torch.equal(h_rew, hr) and torch.equal(h_term, ht) and torch.equal(h_stats, hs)
    assert torch.equal(a.blobs, b.blobs)


def test_step_host_zero_copy_staged_and_pageable_agree():
    """cats_step_host works on mapped page-locked buffers in place (the kernel reads the actions and posts its outputs
    over PCIe); pageable buffers, or the debug switch, go through the staging block.  Same bits every way."""
    import ctypes as C
    from r_mabm_b200 import _native as nat
    B, A = 37, 2
    envs = [_env(B, n_agents=A, seed=67) for _ in range(5)]
    for e in envs: e.run(12)
    g = torch.Generator().manual_seed(5)
    lib = envs[0]._lib
    for step in range(3):
        acts = 0.9 + 0.2 * torch.rand((B, A, 2), dtype=torch.float64, generator=g)
        mask = (torch.rand((B, A), generator=g) > 0.3).to(torch.uint8)
        outs = []
        # 0: pinned, outputs in place (default); 1: pinned, staged both ways; 4: pinned, inputs and outputs in place;
        # 2: pageable caller buffers through the C-ABI; 3: device step
        o = envs[0].step_host(acts.pin_memory(), mask.pin_memory(), want_stats=True)
        outs.append([t.clone() for t in o])
        for env_i, mode in ((1, 0), (4, 3)):
            lib.cats_debug_set_host_zero_copy(mode)
            try:
                o = envs[env_i].step_host(acts.pin_memory(), mask.pin_memory(), want_stats=True)
                outs.append([t.clone() for t in o])
            finally:
                lib.cats_debug_set_host_zero_copy(2)
        e = envs[2]
        e._ensure_host_block()
        h_obs = torch.empty((B, A, 2), dtype=torch.float64); h_rew = torch.empty((B, A), dtype=torch.float64)
        h_term = torch.empty((B,), dtype=torch.uint8); h_stats = torch.empty((B, 1, nat.N_STATS), dtype=torch.float64)
        args = e._args(1, e.flags, want_io=False)
        args.d_actions, args.d_action_mask = acts.data_ptr(), mask.data_ptr()
        args.d_obs, args.d_reward, args.d_terminated = h_obs.data_ptr(), h_rew.data_ptr(), h_term.data_ptr()
        args.d_status, args.d_stats = None, h_stats.data_ptr()
        nat.check(lib.cats_step_host(C.byref(args), e._host_scratch.data_ptr()))
        outs.append([h_obs, h_rew, h_term, h_stats])
        do, dr, dt, ds = envs[3].step(acts.cuda(), mask.cuda(), want_stats=True)
        torch.cuda.synchronize()
        ref = [do.cpu(), dr.cpu(), dt.cpu(), ds.cpu()]
        for k, got in enumerate(outs):
            for x, y in zip(got, ref):
                assert torch.equal(x, y), f"step {step}, path {k}"
    for e in envs:
        assert torch.equal(e.blobs, envs[3].blobs)


def test_full_size_invariants_1024_economies():
    """BASELINE.json config 2 at full width: properties that need no oracle (MODEL.md section 6)."""
    B, W, F, N = 1024, 1000, 100, 20
    env = _env(B, seed=71)
    stats = env.run(60, want_stats=True)
    torch.cuda.synchronize()
    Oc = env.field("Oc")
    emp = (Oc >= 0)
    counts = torch.zeros((B, F + N), dtype=torch.int64, device="cuda")
    counts.scatter_add_(1, Oc.clamp(min=0).long(), emp.long())
    leff = torch.cat([env.field("c_Leff"), env.field("k_Leff")], 1).long()
    assert torch.equal(counts, leff)                                    # links == head-counts
    Q, Y = env.field("c_Q"), env.field("c_Y_prev")
    assert ((Q >= 0) & (Q <= Y * (1 + 1e-12) + 1e-12)).all()
    assert (env.field("PA") >= -1e-9).all() and (env.field("c_deb") >= 0).all() and (env.field("k_deb") >= 0).all()
    assert (env.field("c_P") > 0).all() and (env.field("k_P") > 0).all()
    assert torch.isfinite(env.field("scal")).all() and torch.isfinite(stats).all()
    un = stats[:, :, 2]
    assert ((un >= 0) & (un <= 1)).all()
    # economies are independent: a different batch position gives the same economy (spot check)
    solo = _env(1, seed=71, economy_base=517)
    solo.run(60)
    torch.cuda.synchronize()
    assert torch.equal(solo.blobs[0], env.blobs[517])


def test_reference_shaped_adapters():
    """`Cats` / `CatsLog` return the reference's Python types (cats.py:283-321, 558-626) and agree with
    the oracle on the first observation and reward."""
    from r_mabm_b200.environments import Cats, CatsLog
    for cls, log_mode in ((Cats, False), (CatsLog, True)):
        env = cls(T=6, W=300, F=30, N=6, t_burnin=15, n_agents=2, info_level=3, seed=9)
        obs, info = env.reset(seed=9)
        assert len(obs) == 2 and set(obs[0]) == {"firm_stock", "price_delta"} and isinstance(info, dict)
        o = orc.OracleEconomy(300, 30, 6, seed=9, economy=0)
        for _ in range(15):
            o.step()
        fl = orc.FLAG_LOG if log_mode else 0
        ref0 = o.obs(2, fl)
        assert obs[0]["firm_stock"] == pytest.approx(ref0[0, 0], rel=1e-12, abs=1e-12)
        act = [(0.1, -0.1), ("dummy", "dummy")] if log_mode else [(1.1, 0.9), ("dummy", "dummy")]
        obs, rew, term, trunc, info = env.step(act)
        r_obs, r_rew = o.step(actions=[[act[0][0], act[0][1]], [1.0, 1.0]], mask=[1, 0], flags=fl)
        assert isinstance(rew, list) and isinstance(rew[0], float) and isinstance(term, bool) and isinstance(trunc, bool)
        assert rew[0] == r_rew[0] and rew[1] == r_rew[1]
        assert obs[1]["price_delta"] == r_obs[1, 1]
        for key in ("Y_real", "wb", "Un", "bankruptcy_rate", "avg_price", "agents_sales", "others_sales",
                    "Y_nominal_tot", "gdp_deflator", "inflationRate", "consumption", "totK", "dividendsB",
                    "profitsB", "GB", "deficitGDP", "bonds", "agents_production", "others_equity"):
            assert key in info, key
        assert info["Un"] == o.scal("Un") and len(info["agents_sales"]) == 2
        with pytest.raises(ValueError):
            env.step([(1.0, 1.0)])
        env.reward_type = "rms"
        obs, rew, term, trunc, info = env.step([(1.0, 1.0) if not log_mode else (0.0, 0.0)] * 2)
        assert len(rew) == 2 and 0.0 <= rew[0] <= 1.0
        for _ in range(4):
            obs, rew, term, trunc, info = env.step([("dummy", "dummy")] * 2)
        assert trunc is True
