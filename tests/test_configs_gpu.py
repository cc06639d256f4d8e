"""GPU parity at the sizes BASELINE.json names (configs 2, 3 and 4), through the C-ABI.

Oracle = the in-repo CPU restatement of docs/MODEL.md ("parity unpinned" w.r.t. ABCredit.jl).
Bar: int32 state equal, float64 state and statistics equal to the last bit.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import bits, compare_state, make_oracles

pytestmark = pytest.mark.gpu


def _env(B, **kw):
    from r_mabm_b200.engine import BatchedCats
    return BatchedCats(B, **kw)


def test_config2_1024_economies_1000_periods_sampled_against_the_oracle():
    """BASELINE.json config 2: 1,024 default-size economies x 1,000 periods on one GPU.  The oracle
    runs a sample of the economies (global ids spread over the batch) for the same 1,000 periods;
    state and the statistics of every period must be identical.  Economy-periods are independent of
    batch position, so the sample stands for the batch (test_batch_size_and_sharding_invariance)."""
    B, T = 1024, 1000
    env = _env(B, seed=2026)
    ids = [0, 1, 97, 511, 512, 777, 1000, 1023]
    refs = [orc.OracleEconomy(1000, 100, 20, seed=2026, economy=i) for i in ids]
    stats_ref = orc.run_batch(refs, T, n_threads=min(len(ids), os.cpu_count() or 1), want_stats=True)
    stats = torch.cat([env.run(250, want_stats=True) for _ in range(4)], 1)
    torch.cuda.synchronize()
    g = stats[ids].cpu().numpy()
    assert g.shape == stats_ref.shape == (len(ids), T, 16)
    assert np.array_equal(bits(g), bits(stats_ref))

    class Sub:   # the sampled economies of the batch, presented like a small batch
        W, F, N = env.W, env.F, env.N

        @staticmethod
        def field(name):
            return env.field(name)[ids]
    assert compare_state(Sub, refs, "t=1000") == []
    un = stats[:, -200:, 2]
    assert 0.0 < float(un.mean()) < 0.5 and torch.isfinite(stats).all()


def test_config3_monte_carlo_seeds_64_macro_aggregates_agree():
    """north_star: fresh-RNG runs agree on macro aggregates across 64 seeds.  With the counter-based
    generator the agreement is exact, so the distributions (mean / spread of Y_real, Un,
    inflationRate over 64 seeds) coincide to the last bit as well."""
    W, F, N, T = 300, 30, 6, 120
    seeds = list(range(64))   # one Philox stream per global economy id: 64 independent seeds
    env = _env(64, W=W, F=F, N=N, seed=99)
    stats = env.run(T, want_stats=True)
    torch.cuda.synchronize()
    refs = [orc.OracleEconomy(W, F, N, seed=99, economy=i) for i in seeds]
    stats_ref = orc.run_batch(refs, T, n_threads=os.cpu_count() or 1, want_stats=True)
    g = stats.cpu().numpy()
    for col in (0, 2, 7):   # Y_real, Un, inflationRate
        assert np.array_equal(bits(g[:, :, col]), bits(stats_ref[:, :, col]))
        tail_g, tail_r = g[:, -40:, col].mean(1), stats_ref[:, -40:, col].mean(1)
        assert tail_g.mean() == tail_r.mean() and tail_g.std() == tail_r.std()
    # the seeds really are different economies
    assert len({float(x) for x in g[:, -1, 0]}) == 64


def test_config4_marl_rollout_env_policy_and_td_update_on_device():
    """BASELINE.json config 4 (examples/agent_training.py: CatsLog, 2 learned firms): the batched
    rollout keeps observations, actions and rewards on the device.  Every step is checked against
    the oracle stepping the same economies with the actions the device policy chose, and the Q
    tables against the scalar TD(0) rule (/root/reference/rmabm/agents/qlearners.py:156-166)."""
    from r_mabm_b200.environments import CatsLog
    BL = dict(CatsLog._default_gym_spaces_bounds)
    from r_mabm_b200.rollout import BatchedQPolicy, BatchedRollout
    B, A, W, F, N = 6, 2, 300, 30, 6
    env = _env(B, W=W, F=F, N=N, n_agents=A, log_mode=True, seed=31)
    pol = BatchedQPolicy(B, A, BL, alpha=0.1, gamma=0.99, epsilon=0.5, device="cuda", seed=3)
    ro = BatchedRollout(env, pol, t_burnin=20, T=10)
    obs0 = ro.reset(seed=31).clone()
    refs = make_oracles(B, W, F, N, seed=31)
    for o in refs:
        for _ in range(20):                                   # exactly t_burnin periods, as Cats.reset
            o.step()
    first = [o.obs(A, orc.FLAG_LOG) for o in refs]
    assert np.array_equal(bits(obs0.cpu().numpy()), bits(np.stack(first)))
    launches0 = env._lib.cats_launch_count()
    for t in range(10):
        q_before = pol.Q.clone()
        obs_prev = ro.obs.clone()
        reward, term = ro.step(train=True)
        torch.cuda.synchronize()
        s0, s1, a0, a1 = [x.cpu().numpy() for x in pol._last]
        lo0, hi0 = BL["act_production_factor"]; lo1, hi1 = BL["act_price_factor"]
        acts = ro.last_actions.cpu().numpy()
        assert np.allclose(acts, np.stack([lo0 + a0 * (hi0 - lo0) / 10, lo1 + a1 * (hi1 - lo1) / 10], -1), rtol=0, atol=1e-15)
        for b, o in enumerate(refs):
            r_obs, r_rew = o.step(actions=acts[b], mask=[1] * A, flags=orc.FLAG_LOG)
            assert np.array_equal(bits(ro.obs[b].cpu().numpy()), bits(r_obs)), (t, b)
            assert np.array_equal(bits(reward[b].cpu().numpy()), bits(r_rew)), (t, b)
        # TD(0) on exactly one cell per (economy, agent)
        n0, n1 = [x.cpu().numpy() for x in pol.bin_obs(ro.obs)]
        qb, qa = q_before.cpu().numpy(), pol.Q.cpu().numpy()
        changed = (qb != qa).sum()
        assert changed <= B * A
        for b in range(B):
            for a in range(A):
                cur = qb[b, a, s0[b, a], s1[b, a], a0[b, a], a1[b, a]]
                boot = 0.0 if term[b] else 0.99 * qb[b, a, n0[b, a], n1[b, a]].max()
                want = cur + 0.1 * ((float(reward[b, a]) + boot) - cur)
                assert qa[b, a, s0[b, a], s1[b, a], a0[b, a], a1[b, a]] == pytest.approx(want, rel=1e-15, abs=1e-15)
        del obs_prev
    assert env._lib.cats_launch_count() - launches0 == 10     # one period-kernel launch per rollout step
    assert compare_state(env, refs, "after rollout") == []


def test_rollout_auto_reset_gives_every_economy_its_own_episode_clock():
    """8(f) rank 3: per-economy reset on truncation, on the device.  Economies are knocked out of step
    by one masked reset, then the rollout runs past several truncations; an oracle per economy follows
    the same actions and is reset (new model, episode + 1, t_burnin periods) when its clock runs out."""
    from r_mabm_b200.environments import Cats
    from r_mabm_b200.rollout import BatchedQPolicy, BatchedRollout
    BL = dict(Cats._default_gym_spaces_bounds)
    B, A, W, F, N, T, TB = 5, 2, 300, 30, 6, 4, 6
    env = _env(B, W=W, F=F, N=N, n_agents=A, seed=13)
    pol = BatchedQPolicy(B, A, BL, epsilon=1.0, device="cuda", seed=4)
    ro = BatchedRollout(env, pol, t_burnin=TB, T=T, auto_reset=True)
    ro.reset(seed=13)
    refs = make_oracles(B, W, F, N, seed=13)
    for o in refs:
        for _ in range(TB):
            o.step()
    clock = [0] * B
    # stagger: economies 1 and 3 start a new episode two steps in
    for t in range(11):
        if t == 2:
            m = torch.tensor([0, 1, 0, 1, 0], dtype=torch.bool, device="cuda")
            env.reset_where(m, 0); ro._burnin(active=m); ro.t_b.masked_fill_(m, 0)
            ro.obs = env.observation().clone()
            for b in (1, 3):
                refs[b].reset(); clock[b] = 0
                for _ in range(TB):
                    refs[b].step()
        reward, term = ro.step(train=True)
        torch.cuda.synchronize()
        acts = ro.last_actions.cpu().numpy()
        done = ro.last_done.cpu().numpy()
        for b, o in enumerate(refs):
            r_obs, r_rew = o.step(actions=acts[b], mask=[1] * A, flags=0)
            assert np.array_equal(bits(reward[b].cpu().numpy()), bits(r_rew)), (t, b)
            clock[b] += 1
            assert bool(done[b]) == (clock[b] >= T)
            if clock[b] >= T:
                o.reset(); clock[b] = 0
                for _ in range(TB):
                    o.step()
                r_obs = o.obs(A, 0)
            assert np.array_equal(bits(ro.obs[b].cpu().numpy()), bits(r_obs)), (t, b)
        assert compare_state(env, refs, f"t={t}") == []
    assert ro.episodes.cpu().tolist() == [2, 2, 2, 2, 2]
    ep = env.scalar("episode").cpu().tolist()
    assert ep == [2.0, 3.0, 2.0, 3.0, 2.0]
    # a new episode is a new draw: economy 0 after its reset is not its first episode over again
    first = make_oracles(1, W, F, N, seed=13)[0]
    again = make_oracles(1, W, F, N, seed=13)[0]
    again.reset()
    for _ in range(TB):
        first.step(); again.step()
    assert not np.array_equal(first.i32("Oc"), again.i32("Oc"))


def test_active_mask_leaves_the_other_economies_untouched():
    B, W, F, N = 9, 300, 30, 6
    env = _env(B, W=W, F=F, N=N, seed=2)
    env.run(8)
    before = env.blobs.clone()
    m = torch.tensor([1, 0, 0, 1, 1, 0, 1, 0, 0], dtype=torch.uint8, device="cuda")
    stats = env.run(5, want_stats=True, active=m)
    torch.cuda.synchronize()
    refs = make_oracles(B, W, F, N, seed=2)
    for b, o in enumerate(refs):
        for _ in range(8 + (5 if m[b] else 0)):
            o.step()
    assert compare_state(env, refs, "masked run") == []
    idle = (m == 0).nonzero().flatten().tolist()
    for b in idle:
        assert torch.equal(env.blobs[b], before[b]) and float(stats[b].abs().sum()) == 0.0
    assert float(stats[0, -1, 2]) == refs[0].scal("Un")


def _invariants(env, stats=None):
    """Properties every faithful CATS state has (docs/MODEL.md section 6), on the whole batch."""
    B, F, N = env.B, env.F, env.N
    Oc = env.field("Oc")
    emp = (Oc >= 0)
    counts = torch.zeros((B, F + N), dtype=torch.int64, device="cuda")
    counts.scatter_add_(1, Oc.clamp(min=0).long(), emp.long())
    leff = torch.cat([env.field("c_Leff"), env.field("k_Leff")], 1).long()
    assert torch.equal(counts, leff)                                    # employment links == head-counts
    Q, Y = env.field("c_Q"), env.field("c_Y_prev")
    assert ((Q >= 0) & (Q <= Y * (1 + 1e-12) + 1e-12)).all()            # sales within production
    assert (env.field("PA") >= -1e-9).all() and (env.field("c_deb") >= 0).all() and (env.field("k_deb") >= 0).all()
    assert (env.field("c_P") > 0).all() and (env.field("k_P") > 0).all()
    assert torch.isfinite(env.field("scal")).all()
    # Un is taken by update_tracking_variables! BEFORE firms_go_bankrupt! releases the workers of the
    # firms that exit (cats.py:401-404): it can only be below the end-of-period share of unemployed
    un = env.scalar("Un")
    assert ((un >= 0) & (un <= 1.0 - emp.sum(1).double() / env.W + 1e-12)).all()
    if stats is not None:
        assert torch.isfinite(stats).all()


def test_config3_bench_width_16384_economies_invariants_and_sampled_parity():
    """BASELINE.json config 3 at the width bench.py runs (16,384 default-size economies on one GPU, the
    production kernel): invariants on every economy after the burn-in and 20 more periods, and a sample
    of economies spread over the batch bit-identical to the oracle after the same 320 periods."""
    B, T = 16384, 320
    env = _env(B, seed=20260119)
    env.run(300)
    stats = env.run(20, want_stats=True)
    torch.cuda.synchronize()
    _invariants(env, stats)
    ids = [0, 13, 14, 4143, 4144, 8191, 12345, 16383]          # block and wave boundaries included
    refs = [orc.OracleEconomy(1000, 100, 20, seed=20260119, economy=i) for i in ids]
    stats_ref = orc.run_batch(refs, T, n_threads=min(len(ids), os.cpu_count() or 1), want_stats=True)
    assert np.array_equal(bits(stats[ids].cpu().numpy()), bits(stats_ref[:, 300:]))

    class Sub:
        W, F, N = env.W, env.F, env.N

        @staticmethod
        def field(name):
            return env.field(name)[ids]
    assert compare_state(Sub, refs, "t=320") == []
    # Monte-Carlo seeds are different economies, and the batch does not care how it is split
    assert len({float(x) for x in stats[:, -1, 0].cpu()}) > 16000
    half = _env(8192, seed=20260119, economy_base=8192)
    half.run(300); half.run(20)
    torch.cuda.synchronize()
    assert torch.equal(half.blobs, env.blobs[8192:])


def test_config5_full_width_256_scaled_economies():
    """BASELINE.json config 5 at full width: 256 economies of 10x size.  Invariants on all of them, two of
    them against the oracle."""
    B, W, F, N, T = 256, 10000, 1000, 200, 12
    env = _env(B, W=W, F=F, N=N, seed=5)
    stats = env.run(T, want_stats=True)
    torch.cuda.synchronize()
    _invariants(env, stats)
    ids = [0, 255]
    refs = [orc.OracleEconomy(W, F, N, seed=5, economy=i) for i in ids]
    stats_ref = orc.run_batch(refs, T, n_threads=2, want_stats=True)
    assert np.array_equal(bits(stats[ids].cpu().numpy()), bits(stats_ref))

    class Sub:
        pass
    Sub.W, Sub.F, Sub.N = W, F, N
    Sub.field = staticmethod(lambda name: env.field(name)[ids])
    assert compare_state(Sub, refs, f"t={T}") == []
