"""The batched tabular-Q policy of r_mabm_b200/rollout.py against the reference-shaped scalar QLearner
(/root/reference/rmabm/agents/qlearners.py:63-166; known answers of tests/test_agents.py:48-108).
CPU only: the policy is plain torch, the period kernel is not involved."""
import numpy as np
import pytest
import torch

from r_mabm_b200.agents import QLearner
from r_mabm_b200.rollout import BatchedQPolicy

BOUNDS = {"obs_firm_stock": (-5.0, 5.0), "obs_price_delta": (-2.0, 8.0),
          "act_production_factor": (0.7, 1.3), "act_price_factor": (0.7, 1.3)}


class FakeEnv:
    gym_spaces_bounds = BOUNDS


def _obs_t(x):
    return torch.tensor(x, dtype=torch.float64)


def test_binning_matches_scalar_learner_including_ties_and_clamps():
    pol = BatchedQPolicy(1, 1, BOUNDS, device="cpu")
    ag = QLearner(0, FakeEnv())
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-8, 8, 400), [-5.0, 5.0, 0.0, 0.5, -0.5, 4.5, 100.0, -100.0]])
    ys = np.concatenate([rng.uniform(-5, 11, 400), [-2.0, 8.0, 3.0, 3.5, 2.5, 7.5, -100.0, 100.0]])
    for x, y in zip(xs, ys):
        s0, s1 = pol.bin_obs(_obs_t([[[x, y]]]))
        assert (int(s0), int(s1)) == ag.bin_obs({"firm_stock": x, "price_delta": y}), (x, y)


def test_greedy_action_and_td_update_match_scalar_learner():
    B, A = 3, 2
    pol = BatchedQPolicy(B, A, BOUNDS, alpha=0.1, gamma=0.99, epsilon=0.0, device="cpu", seed=1)
    ags = [[QLearner(a, FakeEnv(), alpha=0.1, gamma=0.99, epsilon_zero=0.0) for a in range(A)] for _ in range(B)]
    rng = np.random.default_rng(11)
    q0 = rng.normal(size=(B, A, 11, 11, 11, 11))
    pol.Q.copy_(torch.from_numpy(q0))
    for b in range(B):
        for a in range(A):
            ags[b][a].Q[...] = q0[b, a]
    obs = np.stack([rng.uniform(-5, 5, (B, A)), rng.uniform(-2, 8, (B, A))], -1)
    for step in range(6):
        act = pol.get_action(torch.from_numpy(obs)).numpy()
        for b in range(B):
            for a in range(A):
                ref = ags[b][a].get_action({"firm_stock": obs[b, a, 0], "price_delta": obs[b, a, 1]})
                assert act[b, a] == pytest.approx(ref, rel=0, abs=1e-15)
        nxt = np.stack([rng.uniform(-5, 5, (B, A)), rng.uniform(-2, 8, (B, A))], -1)
        rew = rng.normal(size=(B, A))
        term = np.array([False, step == 3, False])
        pol.train(torch.from_numpy(rew), torch.from_numpy(nxt), torch.from_numpy(term))
        for b in range(B):
            for a in range(A):
                ags[b][a].train({"firm_stock": obs[b, a, 0], "price_delta": obs[b, a, 1]}, rew[b, a],
                                {"firm_stock": nxt[b, a, 0], "price_delta": nxt[b, a, 1]}, bool(term[b]))
                assert np.allclose(pol.Q[b, a].numpy(), ags[b][a].Q, rtol=0, atol=1e-15)
        obs = nxt


def test_exploration_draws_cover_the_action_grid():
    pol = BatchedQPolicy(64, 2, BOUNDS, epsilon=1.0, device="cpu", seed=5)
    seen0, seen1 = set(), set()
    for _ in range(20):
        act = pol.get_action(torch.zeros((64, 2, 2), dtype=torch.float64))
        assert ((act >= 0.7 - 1e-12) & (act <= 1.3 + 1e-12)).all()
        seen0 |= set(np.round((act[..., 0].numpy() - 0.7) / 0.06).astype(int).ravel().tolist())
        seen1 |= set(np.round((act[..., 1].numpy() - 0.7) / 0.06).astype(int).ravel().tolist())
    assert seen0 == set(range(11)) and seen1 == set(range(11))
