"""`Simulation`: the episode driver behind the reference's examples, over this framework's `Cats`.

Interface of rmabm.Simulation (/root/reference/rmabm/simulation.py:12-207): constructor arguments,
`step` / `run_episode` / `train_agents` / `reset` / `init_logger`, the `*_history` properties, the
epsilon schedule max(0.01, 0.9 ** episode) -- so `examples/heuristic_run.py` and
`examples/agent_training.py` run unchanged.  Thousands of economies at once go through
`r_mabm_b200.rollout.BatchedRollout` instead; this class is the one-economy, host-side loop.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Callable

import numpy as np

from .agents import QLearner
from .logger import Logger


class _Trace:
    """What one episode leaves behind: a row per step of (observations, actions, rewards, info)."""

    __slots__ = ("obs", "act", "reward", "info")

    def __init__(self):
        self.obs, self.act, self.reward, self.info = [], [], [], []

    def add(self, obs, act, reward, info) -> None:
        self.obs.append(obs)
        self.act.append(act)
        self.reward.append(reward)
        self.info.append(info)

    def write(self, logger: Logger, episode: int) -> None:
        # the files rmabm.Logger users read: ep{n}_obs, ep{n}_act, ep{n}_rewards, ep{n}_{info key}
        logger.log_obs(self.obs, episode)
        logger.log_act(self.act, episode)
        logger.log_array(np.array(self.reward), f"ep{episode}_rewards")
        logger.log_info(self.info, episode)


class Simulation:
    def __init__(self, environment, agents: list, n_episodes: int = 100, current_episode: int = 1):
        if len(agents) != environment.n_agents:
            raise ValueError("Number of agents must match number of agents in the environment.")
        self.env, self.agents = environment, agents
        self.qlearners = [ag for ag in agents if isinstance(ag, QLearner)]
        self.n_episodes, self.current_episode = n_episodes, current_episode
        self.logger: Logger | None = None
        self._begin(*self.env.reset())

    # -- episode state --------------------------------------------------------------------------
    def _begin(self, obs, info) -> None:
        self._trace = _Trace()
        self._obs, self._info = obs, info
        self._next_obs = self._reward = None
        self._terminated = self._truncated = False

    @property
    def _done(self) -> bool:
        return bool(self._terminated or self._truncated)

    # the reference keeps four parallel lists under these names (simulation.py:60-64); same data here
    _obs_history = property(lambda self: self._trace.obs)
    _act_history = property(lambda self: self._trace.act)
    _reward_history = property(lambda self: self._trace.reward)
    _info_history = property(lambda self: self._trace.info)
    obs_history, act_history = _obs_history, _act_history
    reward_history, info_history = _reward_history, _info_history

    # -- public interface -------------------------------------------------------------------------
    def init_logger(self, log_name: str, log_directory: str | Path, use_timestamp: bool = True) -> None:
        self.logger = Logger(log_name, log_directory, use_timestamp)

    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None) -> None:
        """New model + burn-in in the environment (which also seeds `random` and numpy for the agents)."""
        self._begin(*self.env.reset(seed=seed, options=options))

    def step(self, train: bool = False) -> None:
        """Agents act on the current observations, the economy advances one period, and -- if asked --
        every agent learns from its own (observation, reward, next observation)."""
        seen = self._obs
        chosen = [ag.get_action(o) for ag, o in zip(self.agents, seen)]
        self._next_obs, self._reward, self._terminated, self._truncated, self._info = self.env.step(chosen)
        self._trace.add(seen, chosen, self._reward, self._info)
        if train:
            for k, ag in enumerate(self.agents):
                ag.train(seen[k], self._reward[k], self._next_obs[k], self._terminated)
        self._obs = self._next_obs

    def run_episode(self, train: bool = False) -> None:
        """Step until the episode terminates or is truncated, write the log (if a logger is set), then
        reset for the next episode."""
        while not self._done:
            self.step(train=train)
        if self.logger is not None:
            self._trace.write(self.logger, self.current_episode)
        self.reset()
        self.current_episode += 1

    def train_agents(self, n_episodes: int | None = None, update: Callable[..., None] | None = None) -> None:
        """Training episodes; after each one `update(agent, simulation)` is applied to the Q-learners
        (default: the reference's exploration schedule)."""
        todo = n_episodes or (self.n_episodes - self.current_episode + 1)
        schedule = update or self._default_update
        for _ in range(todo):
            self.run_episode(train=True)
            for learner in self.qlearners:
                schedule(learner, self)

    @staticmethod
    def _default_update(agent: QLearner, simulation: "Simulation") -> None:
        agent.epsilon = max(0.01, 0.9 ** simulation.current_episode)
