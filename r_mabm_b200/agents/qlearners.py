"""Tabular Q-learner with the interface and numerics of rmabm.agents.QLearner
(/root/reference/rmabm/agents/qlearners.py:8-166): nearest-pole binning of the two observations,
epsilon-greedy over an n_bins[2] x n_bins[3] action grid (global numpy RNG, as the reference),
TD(0) update without the bootstrap term on termination.  The device-resident, batched version of
the same policy lives in r_mabm_b200/rollout.py.
"""
from __future__ import annotations

import warnings

import numpy as np


class QLearner:
    def __init__(self, agent_id: int, environment, n_bins=(11, 11, 11, 11), alpha: float = 0.1,
                 gamma: float = 0.99, epsilon_zero: float = 0.9):
        self.agent_id = agent_id
        self.alpha, self.gamma, self.epsilon = alpha, gamma, epsilon_zero
        if isinstance(n_bins, int):
            n_bins = (n_bins,) * 4
        self.Q = np.zeros(shape=n_bins)
        self.bounds = environment.gym_spaces_bounds
        if not 0 < alpha < 1:
            warnings.warn("alpha should be in (0, 1)")
        if not 0 < gamma < 1:
            warnings.warn("gamma should be in (0, 1)")
        if not 0 <= epsilon_zero <= 1:
            warnings.warn("epsilon should be in [0, 1]")
        self._last_action = None

    def _poles(self, key: str, n: int) -> np.ndarray:
        lo, hi = self.bounds[key]
        return np.linspace(lo, hi, n)

    def bin_obs(self, obs: dict) -> tuple[int, int]:
        """Index of the closest pole for each observation (first index wins ties; clamps outside)."""
        i = int(np.abs(self._poles("obs_firm_stock", self.Q.shape[0]) - obs["firm_stock"]).argmin())
        j = int(np.abs(self._poles("obs_price_delta", self.Q.shape[1]) - obs["price_delta"]).argmin())
        return i, j

    def get_action(self, obs: dict) -> tuple[float, float]:
        if np.random.rand() < self.epsilon:
            a0 = np.random.randint(0, self.Q.shape[2])
            a1 = np.random.randint(0, self.Q.shape[3])
        else:
            i, j = self.bin_obs(obs)
            a0, a1 = np.unravel_index(np.argmax(self.Q[i, j]), self.Q.shape[2:])
        self._last_action = (a0, a1)
        lo0, hi0 = self.bounds["act_production_factor"]
        lo1, hi1 = self.bounds["act_price_factor"]
        return (lo0 + a0 * (hi0 - lo0) / (self.Q.shape[2] - 1),
                lo1 + a1 * (hi1 - lo1) / (self.Q.shape[3] - 1))

    def train(self, obs: dict, reward: float, next_obs: dict | None, terminated: bool) -> None:
        if next_obs is None and not terminated:
            raise TypeError("next_obs can't be None if the episode is not terminated")
        idx = (*self.bin_obs(obs), *self._last_action)
        if terminated:
            delta = reward - self.Q[idx]
        else:
            i, j = self.bin_obs(next_obs)
            delta = reward + self.gamma * np.max(self.Q[i, j]) - self.Q[idx]
        self.Q[idx] += self.alpha * delta
