"""GPU: the fused tabular-Q policy kernel (`cats_q_step`: TD update + epsilon-greedy choice in one launch,
include/cats_b200.h) against a numpy restatement of rmabm.agents.QLearner
(/root/reference/rmabm/agents/qlearners.py:63-166) fed the same Philox exploration draws, and against the
reference-shaped scalar `QLearner` for the update rule.  Q tables, bins and actions are compared to the bit."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import bits

pytestmark = pytest.mark.gpu


def philox(c0, c1, c2, c3, seed):
    out = (C.c_uint32 * 4)()
    orc.lib().oracle_philox(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3),
                            C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32), out)
    return [int(x) for x in out]


class NumpyQ:
    """qlearners.py in numpy, one table per (economy, agent), exploration from the kernel's Philox stream."""

    def __init__(self, B, A, bounds, n, alpha, gamma, eps, seed, economy_base):
        self.B, self.A, self.b, self.n = B, A, bounds, n
        self.alpha, self.gamma, self.eps, self.seed, self.base = alpha, gamma, eps, seed, economy_base
        self.Q = np.zeros((B, A) + tuple(n))
        self.last = np.zeros((B, A, 4), dtype=np.int64)
        self.step = 0

    def bin(self, x, key, n):
        lo, hi = self.b[key]
        step = (hi - lo) / (n - 1)
        idx = np.ceil((x - lo) / step - 0.5)
        if idx != idx:
            idx = 0.0
        return int(min(max(idx, 0.0), n - 1))

    def train(self, reward, obs, term):
        for b in range(self.B):
            for a in range(self.A):
                s0, s1, a0, a1 = self.last[b, a]
                n0, n1 = self.bin(obs[b, a, 0], "obs_firm_stock", self.n[0]), self.bin(obs[b, a, 1], "obs_price_delta", self.n[1])
                boot = 0.0 if term[b] else self.gamma * self.Q[b, a, n0, n1].max()
                cur = self.Q[b, a, s0, s1, a0, a1]
                self.Q[b, a, s0, s1, a0, a1] = cur + self.alpha * ((reward[b, a] + boot) - cur)

    def act(self, obs):
        out = np.zeros((self.B, self.A, 2))
        for b in range(self.B):
            for a in range(self.A):
                s0, s1 = self.bin(obs[b, a, 0], "obs_firm_stock", self.n[0]), self.bin(obs[b, a, 1], "obs_price_delta", self.n[1])
                a0, a1 = np.unravel_index(np.argmax(self.Q[b, a, s0, s1]), self.n[2:])
                w = philox(a, 0x51, self.step, self.base + b, self.seed)
                if w[0] * (1.0 / 4294967296.0) < self.eps:
                    a0, a1 = (w[1] * self.n[2]) >> 32, (w[2] * self.n[3]) >> 32
                self.last[b, a] = (s0, s1, a0, a1)
                lo0, hi0 = self.b["act_production_factor"]; lo1, hi1 = self.b["act_price_factor"]
                out[b, a] = (lo0 + a0 * (hi0 - lo0) / (self.n[2] - 1), lo1 + a1 * (hi1 - lo1) / (self.n[3] - 1))
        self.step += 1
        return out


@pytest.mark.parametrize("log_mode", [False, True])
def test_fused_policy_rollout_matches_the_numpy_learners(log_mode):
    from r_mabm_b200.engine import BatchedCats
    from r_mabm_b200.environments import Cats, CatsLog
    from r_mabm_b200.rollout import BatchedRollout, FusedQPolicy
    bounds = dict((CatsLog if log_mode else Cats)._default_gym_spaces_bounds)
    B, A, n = 5, 2, (5, 4, 3, 3)
    env = BatchedCats(B, W=300, F=30, N=6, n_agents=A, log_mode=log_mode, seed=8, economy_base=40)
    pol = FusedQPolicy(B, A, bounds, n_bins=n, alpha=0.2, gamma=0.9, epsilon=0.4, seed=77, economy_base=40)
    ref = NumpyQ(B, A, bounds, n, 0.2, 0.9, 0.4, 77, 40)
    ro = BatchedRollout(env, pol, t_burnin=15, T=1000)
    obs = ro.reset(seed=8).cpu().numpy()
    want = ref.act(obs)
    launches0 = env._lib.cats_launch_count()
    for t in range(30):
        reward, term = ro.step(train=True)
        torch.cuda.synchronize()
        assert np.array_equal(bits(ro.last_actions.cpu().numpy()), bits(want)), t
        obs, rew, tm = ro.obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy()
        ref.train(rew, obs, tm)
        assert np.array_equal(bits(pol.Q.cpu().numpy()), bits(ref.Q)), t
        want = ref.act(obs)
        assert np.array_equal(pol.last.cpu().numpy(), ref.last), t
    assert env._lib.cats_launch_count() - launches0 == 1 + 2 * 30      # first action, then period + policy kernel per step
    assert np.count_nonzero(ref.Q) > 20


def test_fused_update_is_the_scalar_qlearner_rule_and_masks_work():
    from r_mabm_b200.agents import QLearner
    from r_mabm_b200.environments import Cats
    from r_mabm_b200.rollout import FusedQPolicy
    bounds = dict(Cats._default_gym_spaces_bounds)
    B, A = 4, 3
    pol = FusedQPolicy(B, A, bounds, n_bins=(11, 11, 11, 11), alpha=0.1, gamma=0.99, epsilon=0.0, seed=1)
    rng = np.random.default_rng(0)
    pol.Q.copy_(torch.from_numpy(rng.normal(size=tuple(pol.Q.shape))))
    q0 = pol.Q.cpu().numpy().copy()
    obs = torch.from_numpy(rng.uniform(-6, 9, size=(B, A, 2))).cuda()      # some outside the bounds: clamps
    obs[0, 0, 0] = float("nan")
    acts = pol.get_action(obs).cpu().numpy()
    last = pol.last.cpu().numpy().copy()
    nxt = torch.from_numpy(rng.uniform(-6, 9, size=(B, A, 2))).cuda()
    rew = torch.from_numpy(rng.normal(size=(B, A))).cuda()
    term = torch.tensor([0, 1, 0, 0], dtype=torch.uint8, device="cuda")
    active = torch.tensor([1, 1, 0, 1], dtype=torch.uint8, device="cuda")
    pol._launch(nxt, rew, term, True, False, active)
    torch.cuda.synchronize()
    q1 = pol.Q.cpu().numpy()

    class E:
        gym_spaces_bounds = bounds
    for b in range(B):
        for a in range(A):
            ql = QLearner(a, E(), n_bins=(11, 11, 11, 11), alpha=0.1, gamma=0.99, epsilon_zero=0.0)
            ql.Q = q0[b, a].copy()
            o = {"firm_stock": float(obs[b, a, 0]), "price_delta": float(obs[b, a, 1])}
            if b == 0 and a == 0:
                o["firm_stock"] = bounds["obs_firm_stock"][0]        # the kernel bins NaN to 0
            got = ql.get_action(o)                                    # epsilon 0: greedy
            assert tuple(last[b, a][:2]) == tuple(int(x) for x in ql.bin_obs(o))
            assert tuple(last[b, a][2:]) == tuple(int(x) for x in ql._last_action)
            assert acts[b, a, 0] == got[0] and acts[b, a, 1] == got[1]
            ql.train(o, float(rew[b, a]), {"firm_stock": float(nxt[b, a, 0]), "price_delta": float(nxt[b, a, 1])}, bool(term[b]))
            if active[b]:
                np.testing.assert_allclose(q1[b, a], ql.Q, rtol=0, atol=1e-15)
                assert (q1[b, a] != q0[b, a]).sum() <= 1
            else:
                assert np.array_equal(q1[b, a], q0[b, a])


def test_fused_rollout_does_not_depend_on_sharding():
    """Economies 0..5 as one batch, and as two shards with `economy_base` (what two ranks would run): same
    actions, rewards and Q tables -- env draws and exploration draws are keyed by the global economy id."""
    from r_mabm_b200.engine import BatchedCats
    from r_mabm_b200.environments import CatsLog
    from r_mabm_b200.rollout import BatchedRollout, FusedQPolicy
    bounds = dict(CatsLog._default_gym_spaces_bounds)

    def run(B, base):
        env = BatchedCats(B, W=300, F=30, N=6, n_agents=2, log_mode=True, seed=4, economy_base=base)
        pol = FusedQPolicy(B, 2, bounds, n_bins=(5, 5, 3, 3), epsilon=0.5, seed=9, economy_base=base)
        ro = BatchedRollout(env, pol, t_burnin=10, T=100)
        ro.reset(seed=4)
        rewards = torch.stack([ro.step(train=True)[0].clone() for _ in range(15)])
        torch.cuda.synchronize()
        return rewards.cpu(), pol.Q.cpu(), env.blobs.cpu()
    full = run(6, 0)
    lo, hi = run(2, 0), run(4, 2)
    assert torch.equal(full[0], torch.cat([lo[0], hi[0]], 1))
    assert torch.equal(full[1], torch.cat([lo[1], hi[1]], 0))
    assert torch.equal(full[2], torch.cat([lo[2], hi[2]], 0))
