"""`Dummy` agent: leaves the firm's heuristic decision untouched.

Interface of rmabm.agents.Dummy (/root/reference/rmabm/agents/heuristic.py:6-80): `agent_id`,
`action_length`, `get_action()` -> a tuple of "dummy" markers that the environment ignores
(cats.py:439-440), `bin_obs()` -> (-1, -1), `train()` -> no-op.  In the batched engine a Dummy is a
False entry of the action mask.
"""
from __future__ import annotations


class Dummy:
    def __init__(self, agent_id: int, environment=None):
        self.agent_id = agent_id
        self.action_length: int = 2
        if environment:
            shape = getattr(environment.action_space, "shape", None)
            try:
                self.action_length = shape[0]
            except TypeError:  # a Dict action space has no shape
                self.action_length = len(environment.action_space.keys())

    @staticmethod
    def bin_obs(*args, **kwargs) -> tuple[int, int]:
        return -1, -1

    def get_action(self, *args, **kwargs) -> tuple[str, ...]:
        return ("dummy",) * self.action_length

    @staticmethod
    def train(*args, **kwargs) -> None:
        return None
