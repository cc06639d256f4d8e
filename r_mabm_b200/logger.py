"""`Logger`: the reference's on-disk format, unchanged (/root/reference/rmabm/logger.py:9-202).

One `.npy` per info key / obs / act / rewards under `log_directory/log_name[_timestamp]/`, files
named `ep{n}_{key}.npy`, append-by-reload when a file exists.  `log_batched_stats` additionally
writes the engine's [B, T, 16] statistics tensor in the same per-key layout for many economies.
"""
from __future__ import annotations

from datetime import datetime
from pathlib import Path
from typing import Any

import numpy as np


class Logger:
    def __init__(self, log_name: str, log_directory: Path | str, use_timestamp: bool = True):
        self._log_name = log_name
        if use_timestamp:
            self._log_name = f"{log_name}_{datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}"
        self.path = Path(log_directory) / self._log_name
        self.path.mkdir(parents=True, exist_ok=True)

    def log_array(self, array: np.ndarray, filename: str) -> None:
        if array.size == 0:                       # no empty files
            return
        if not filename.endswith(".npy"):
            filename += ".npy"
        target = self.path / filename
        if target.exists():
            array = np.concatenate([np.load(target), array], axis=0)
        np.save(target, array)

    def log_info(self, info_list: list[dict[str, Any]], episode: int | None = None) -> None:
        if not info_list or not info_list[0] or not info_list[0].keys():
            return
        for key in info_list[0].keys():
            values = np.array([d[key] for d in info_list])
            self.log_array(values, f"ep{episode}_{key}" if episode is not None else str(key))

    def _log_gym(self, gym_list: list[list], filename: str) -> None:
        if not gym_list or not gym_list[0]:
            return
        first = gym_list[0][0]
        if isinstance(first, dict):
            dim, kind = len(first.keys()), "obs"
        elif isinstance(first, tuple):
            dim, kind = len(first), "act"
        else:
            raise TypeError("The gym_list must contain dictionaries or tuples.")
        out = np.empty((len(gym_list), len(gym_list[0]), dim))
        for t, per_agent in enumerate(gym_list):
            for a, item in enumerate(per_agent):
                if "dummy" in item:               # dummy actions are logged as zeros
                    out[t, a, :] = 0
                else:
                    out[t, a, :] = list(item.values()) if kind == "obs" else item
        self.log_array(out, filename)

    def log_act(self, act_list: list[list], episode: int | None = None) -> None:
        if all("dummy" in act for act in act_list[0]):
            return
        self._log_gym(act_list, "act.npy" if episode is None else f"ep{episode}_act.npy")

    def log_obs(self, obs_list: list[list], episode: int | None = None) -> None:
        self._log_gym(obs_list, "obs.npy" if episode is None else f"ep{episode}_obs.npy")

    # -- batched engine output --------------------------------------------------------------------
    def log_batched_stats(self, stats: np.ndarray, episode: int | None = None, keys=None) -> None:
        """stats: [B, T, 16] as returned by BatchedCats.run(want_stats=True).  Writes one file per
        aggregate with shape (T, B): the reference's per-key layout with an economy axis appended."""
        from ._native import STAT_NAMES
        for i, key in enumerate(STAT_NAMES):
            if keys is not None and key not in keys:
                continue
            self.log_array(np.ascontiguousarray(stats[:, :, i].T), f"ep{episode}_{key}" if episode is not None else key)
