from .qlearners import QLearner
from .heuristic import Dummy

__all__ = ["QLearner", "Dummy"]
