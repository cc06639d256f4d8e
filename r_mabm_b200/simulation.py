"""`Simulation`: episode driver with the interface of rmabm.Simulation
(/root/reference/rmabm/simulation.py:12-207): `step`, `run_episode`, `train_agents`, `reset`,
`init_logger`, `*_history` properties, epsilon schedule max(0.01, 0.9 ** episode).
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Callable

import numpy as np

from .agents import Dummy, QLearner
from .logger import Logger


class Simulation:
    def __init__(self, environment, agents: list, n_episodes: int = 100, current_episode: int = 1):
        if environment.n_agents != len(agents):
            raise ValueError("Number of agents must match number of agents in the environment.")
        self.env = environment
        self.agents = agents
        self.qlearners = [a for a in agents if isinstance(a, QLearner)]
        self.n_episodes = n_episodes
        self.current_episode = current_episode
        self.logger: Logger | None = None
        self._clear_histories()
        self._obs, self._info = self.env.reset()
        self._next_obs = self._reward = None
        self._terminated = self._truncated = False

    def _clear_histories(self) -> None:
        self._obs_history, self._act_history, self._reward_history, self._info_history = [], [], [], []

    def init_logger(self, log_name: str, log_directory: str | Path, use_timestamp: bool = True) -> None:
        self.logger = Logger(log_name, log_directory, use_timestamp)

    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None) -> None:
        self._obs, self._info = self.env.reset(seed=seed, options=options)
        self._next_obs = self._reward = None
        self._terminated = self._truncated = False
        self._clear_histories()

    def step(self, train: bool = False) -> None:
        actions = [agent.get_action(obs) for agent, obs in zip(self.agents, self._obs)]
        self._next_obs, self._reward, self._terminated, self._truncated, self._info = self.env.step(actions)
        self._obs_history.append(self._obs)
        self._act_history.append(actions)
        self._reward_history.append(self._reward)
        self._info_history.append(self._info)
        if train:
            for i, agent in enumerate(self.agents):
                agent.train(self._obs[i], self._reward[i], self._next_obs[i], self._terminated)
        self._obs = self._next_obs

    def run_episode(self, train: bool = False) -> None:
        while not self._terminated and not self._truncated:
            self.step(train=train)
        if self.logger:
            self.logger.log_obs(self._obs_history, self.current_episode)
            self.logger.log_act(self._act_history, self.current_episode)
            self.logger.log_array(np.array(self._reward_history), f"ep{self.current_episode}_rewards")
            self.logger.log_info(self._info_history, self.current_episode)
        self.reset()
        self.current_episode += 1

    def train_agents(self, n_episodes: int | None = None, update: Callable[..., None] | None = None) -> None:
        n_episodes = n_episodes or (self.n_episodes - self.current_episode + 1)
        update = update or self._default_update
        for _ in range(n_episodes):
            self.run_episode(train=True)
            for agent in self.qlearners:
                update(agent, self)

    @staticmethod
    def _default_update(agent: QLearner, simulation) -> None:
        agent.epsilon = max(0.01, 0.9 ** simulation.current_episode)

    @property
    def obs_history(self):
        return self._obs_history

    @property
    def act_history(self):
        return self._act_history

    @property
    def reward_history(self):
        return self._reward_history

    @property
    def info_history(self):
        return self._info_history
