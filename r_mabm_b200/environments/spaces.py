"""Minimal stand-ins for gymnasium.spaces.Box / Dict.

gymnasium is an optional dependency here (it is not installed in the build image).  The reference
builds `observation_space` / `action_space` as `gym.spaces.Dict` of scalar `Box`es
(/root/reference/rmabm/environments/cats.py:259-278); when gymnasium is importable those classes are
used, otherwise these stand-ins provide what the reference's callers touch: `.low`, `.high`,
`.keys()`, item access, `.sample()`, `.contains()` (`Dummy` relies on `len(action_space.keys())`,
/root/reference/rmabm/agents/heuristic.py:28-33).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium exists
    import gymnasium as _gym
    Box = _gym.spaces.Box
    Dict = _gym.spaces.Dict
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent: tiny local equivalents
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.full(shape or (1,), low, dtype=dtype)
            self.high = np.full(shape or (1,), high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x, dtype=np.float64).reshape(self.shape)
            return bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.item(0)}, {self.high.item(0)}, {self.shape}, {self.dtype})"

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)
            self.shape = None  # as gymnasium: a Dict space has no shape

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def values(self):
            return self.spaces.values()

        def __getitem__(self, key):
            return self.spaces[key]

        def __len__(self):
            return len(self.spaces)

        def __iter__(self):
            return iter(self.spaces)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x) -> bool:
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"
