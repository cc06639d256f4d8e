"""`Cats` / `CatsLog`: the reference's gymnasium-style environment API over the B200 engine.

Same constructor arguments, attributes, `reset` / `step` signatures, return types, info keys and
ValueErrors as `rmabm.environments.Cats` / `CatsLog`
(/root/reference/rmabm/environments/cats.py:20-704).  What changed is below the boundary: instead
of ~26 juliacall calls into ABCredit.jl per step, one launch of the CUDA period kernel advances the
economy (r_mabm_b200/engine.py, include/cats_b200.h).  This adapter is the B = 1 view of
`BatchedCats`; RL rollouts over thousands of economies should use `BatchedCats` /
`r_mabm_b200.rollout` directly and keep actions and observations on the device.
"""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, Dict, Tuple

import numpy as np

from ..params import DEFAULT_PARAMETERS, load_parameters
from . import spaces


class Cats:
    """CATS model as a multi-agent gym-style environment (see the reference docstring, cats.py:21-62)."""

    metadata: Dict = {
        "render_modes": ["print", "plot", "save"],
        "render_fps": 4,
    }

    _default_parameters: Dict = dict(DEFAULT_PARAMETERS)

    _default_gym_spaces_bounds: Dict = {
        "obs_firm_stock": (-5.0, 5.0),
        "obs_price_delta": (-2.0, 8.0),
        "act_production_factor": (0.7, 1.3),
        "act_price_factor": (0.7, 1.3),
    }

    _log_mode = False

    def __init__(
            self,
            T: int = 5000,
            W: int = 1000,
            F: int = 100,
            N: int = 20,
            params: Dict[str, Any] | str | Path | None = None,
            t_burnin: int = 300,
            n_agents: int = 0,
            reward_type: str = "profits",
            price_change: str = "agent_price",
            bankruptcy_reward: int = -100,
            render_mode: Dict[str, Any] | None = None,
            gym_spaces_bounds: Dict[str, Tuple[float, float]] | None = None,
            info_level: int = 1,
            seed: int | None = None,
            device: str = "cuda",
    ):
        self.T, self.W, self.F, self.N = T, W, F, N
        self.t_burnin = t_burnin
        self.n_agents = n_agents
        self.bankruptcy_reward = bankruptcy_reward
        self.info_level = info_level

        self.params = load_parameters(params)  # str keys, caller's dict filled in place (cats.py:182-184)

        if reward_type not in ["profits", "rms"]:
            raise ValueError("Invalid reward type. Must be 'profits' or 'rms'")
        self.reward_type = reward_type
        if price_change not in ["agent_price", "avg_price"]:
            raise ValueError("Invalid price change mechanism. Must be 'agent_price' or 'avg_price'")
        self._price_change = price_change
        if render_mode not in self.metadata["render_modes"] and render_mode is not None:
            raise ValueError(f"Invalid render mode. Must be one of {self.metadata['render_modes']}")
        self.render_mode = render_mode

        # the model: one economy on the device (cats.py:164 `_julia_model_init`)
        from ..engine import BatchedCats
        self._seed = 0 if seed is None else int(seed)
        self._engine = BatchedCats(1, W=W, F=F, N=N, params=self.params, n_agents=n_agents,
                                   reward_type=reward_type, price_change=price_change,
                                   log_mode=self._log_mode, bankruptcy_reward=float(bankruptcy_reward),
                                   seed=self._seed, device=device)
        self._alloc_host()
        self._sync_host()
        # init_price is inserted by initialise_model in the reference (read at cats.py:530)
        self.params["init_price"] = self._scal("init_price")

        ids_c_firms = [f + 1 for f in range(F)]          # firm_id: 1-based slot index
        self.agents_ids: list[int] = ids_c_firms[:n_agents]
        self._other_firms_ids = ids_c_firms[n_agents:]

        self._create_spaces(gym_spaces_bounds)
        self.t: int = 0

    # -- spaces (cats.py:245-278) -------------------------------------------------------------
    def _create_spaces(self, gym_spaces_bounds) -> None:
        if gym_spaces_bounds is None:
            gym_spaces_bounds = dict(self._default_gym_spaces_bounds)
        for key, value in self._default_gym_spaces_bounds.items():
            if key not in gym_spaces_bounds:
                gym_spaces_bounds[key] = value
        self.gym_spaces_bounds = gym_spaces_bounds
        b = gym_spaces_bounds
        self.observation_space = spaces.Dict({
            "firm_stock": spaces.Box(low=b["obs_firm_stock"][0], high=b["obs_firm_stock"][1]),
            "price_delta": spaces.Box(low=b["obs_price_delta"][0], high=b["obs_price_delta"][1]),
        })
        self.action_space = spaces.Dict({
            "production_factor": spaces.Box(low=b["act_production_factor"][0], high=b["act_production_factor"][1]),
            "price_factor": spaces.Box(low=b["act_price_factor"][0], high=b["act_price_factor"][1]),
        })

    def render(self):
        pass

    # -- reward_type may be reassigned at run time (reference tests/test_env.py:172) -----------
    @property
    def reward_type(self) -> str:
        return self._reward_type

    @reward_type.setter
    def reward_type(self, value: str) -> None:
        self._reward_type = value
        eng = getattr(self, "_engine", None)
        if eng is not None:
            if value not in ("profits", "rms"):
                raise ValueError(f"Invalid reward type. Must be 'profits' or 'rms'. Got {value}.")
            eng.reward_type = value

    # -- reset / step ---------------------------------------------------------------------------
    def reset(self, seed: int | None = None, options: Dict[str, Any] | None = None):
        """New model + burn-in (cats.py:283-315).  The seed keys the engine's Philox stream; Python's
        `random` and numpy's global RNG are seeded too, as the reference does for the agents."""
        random.seed(seed)
        np.random.seed(seed)
        if seed is not None:
            self._seed = int(seed)
        else:
            self._seed = random.getrandbits(63)          # unseeded reset: a fresh stream, like the reference
        self._engine.reset_state(seed=self._seed)
        self._burnin()
        self._sync_host()
        return self._get_obs(), self._get_info()

    def _burnin(self) -> None:
        if self.t_burnin > 0:
            self._engine.run(self.t_burnin)               # one launch, state stays on chip
        self.t = 0                                         # agg.timestep is NOT reset (cats.py:432)

    def step(self, actions, burnin: bool = False):
        if len(actions) != self.n_agents:
            raise ValueError(
                f"Number of actions ({len(actions)}) must be equal to the number of agents ({self.n_agents})")
        eng = self._engine
        h_act = h_mask = None
        if self.n_agents > 0 and not burnin:
            h_act, h_mask = eng.host_action_buffers()      # pinned views the library uploads with one copy
            vals, mask = h_act.numpy(), h_mask.numpy()
            vals[...] = 1.0; mask[...] = 0
            for i, a in enumerate(actions):
                if "dummy" in a:                           # cats.py:439-440
                    continue
                vals[0, i, 0], vals[0, i, 1] = float(a[0]), float(a[1])
                mask[0, i] = 1
        # ONE call, ONE synchronisation: the kernel writes obs / reward / terminated / status straight into pinned
        # host memory (cats_step_host); the state mirror behind info / attributes is refreshed only when read
        self._h_obs, self._h_rew, self._h_term, _ = eng.step_host(h_act, h_mask, want_stats=False)
        self._h_status = eng._h_status
        self._blob_stale = True
        self.t += 1
        status = int(self._h_status[0])
        if status:
            import warnings
            from .. import _native as nat
            if status & nat.STATUS_DEPVAL_DIV0:
                warnings.warn("division by zero in 'depreciation value', setting to 0.")
            if status & nat.STATUS_PROFIT_NAN:
                warnings.warn("profits are nan (likely underflow), setting to 0.")
        n = self.n_agents
        reward = [float(x) for x in self._h_rew[0, :n].tolist()] if n else []
        terminated = bool(int(self._h_term[0]))
        obs = [{"firm_stock": v[0], "price_delta": v[1]} for v in self._h_obs[0, :n].tolist()] if n else []
        info = self._get_info()
        truncated = self.t >= self.T
        return obs, reward, terminated, truncated, info

    # -- host mirror of the one economy: filled once per step / reset, read by obs / info ---------------
    def _alloc_host(self) -> None:
        import torch
        eng = self._engine
        self._h_blob = torch.empty((eng.blob_bytes,), dtype=torch.uint8, pin_memory=True)
        self._h_views: dict = {}
        self._blob_stale = True

    def _sync_host(self) -> None:
        """Refresh the host mirror of the economy's state (one 37 KB copy at the default size)."""
        import torch
        eng = self._engine
        self._h_blob.copy_(eng.blobs[0], non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        self._blob_stale = False

    def _field(self, name):
        """Host view of a named state field of the economy, as of the last step / reset."""
        if self._blob_stale:
            self._sync_host()
        v = self._h_views.get(name)
        if v is None:
            from .. import _native as nat
            off, cnt, dt = nat.field_info(self.W, self.F, self.N, name)
            st = nat.field_stride(name)
            raw = self._h_blob.numpy()
            arr = raw.view(np.float64 if dt == 0 else np.int32)
            esz = 8 if dt == 0 else 4
            v = self._h_views[name] = arr[off // esz: off // esz + cnt * st: st]
        return v

    def _scal(self, name) -> float:
        from ..engine import SCAL_INDEX
        return float(self._field("scal")[SCAL_INDEX[name]])

    def _get_obs(self):
        """cats.py:466-481 (Cats) / 655-671 (CatsLog), from the current state."""
        out = []
        if self.n_agents == 0:
            return out
        yp, yd, p = self._field("c_Y_prev"), self._field("c_Yd"), self._field("c_P")
        ap = self._scal("price")
        for i in range(self.n_agents):
            out.append(self._make_obs(float(yp[i]), float(yd[i]), float(p[i]), ap))
        return out

    @staticmethod
    def _make_obs(y_prev, yd, p, avg_price):
        return {"firm_stock": y_prev - yd, "price_delta": p - avg_price}

    def _get_terminated(self) -> bool:
        return float(self._field("c_A").sum()) == 0

    def _get_truncated(self) -> bool:
        return self.t >= self.T

    def _get_info(self) -> dict:
        """Same keys and shapes as cats.py:558-626."""
        if self.info_level == 0:
            return {}
        n = self.n_agents
        sc = self._field("scal")
        from ..engine import SCAL_INDEX as S
        Q = self._field("c_Q")
        info = {
            "Y_real": float(sc[S["Y_real"]]), "wb": float(sc[S["wb"]]), "Un": float(sc[S["Un"]]),
            "bankruptcy_rate": float(sc[S["bankruptcy_rate"]]), "avg_price": float(sc[S["price"]]),
            "agents_sales": [float(x) for x in Q[:n]],
            "others_sales": (np.mean(Q[n:]), np.std(Q[n:])),
        }
        if self.info_level >= 2:
            price = float(sc[S["price"]])

            def pair(key, arr, cast=float):
                info["agents_" + key] = [cast(x) for x in arr[:n]]
                info["others_" + key] = (np.mean(arr[n:]), np.std(arr[n:]))

            info.update({
                "Y_nominal_tot": float(sc[S["Y_nominal_tot"]]), "gdp_deflator": float(sc[S["gdp_deflator"]]),
                "inflationRate": float(sc[S["inflationRate"]]), "consumption": float(sc[S["consumption"]]),
                "totK": float(sc[S["totK"]]), "dividendsB": float(sc[S["dividendsB"]]),
                "profitsB": float(sc[S["profitsB"]]), "GB": float(sc[S["GB"]]),
                "deficitGDP": float(sc[S["deficitGDP"]]), "bonds": float(sc[S["gov_bonds"]]),
            })
            pair("production", self._field("c_Y_prev") * price)
            pair("debt", self._field("c_deb"))
            pair("employment", self._field("c_Leff"), int)
            pair("capital", self._field("c_K"))
            pair("equity", self._field("c_A"))
            pair("investment", self._field("c_investment"))
            pair("liquidity", self._field("c_liquidity"))
        return info


class CatsLog(Cats):
    """Cats with logarithmic observations and actions (cats.py:629-704)."""

    _default_gym_spaces_bounds: Dict = {
        "obs_firm_stock": (-2.0, 2.0),
        "obs_price_delta": (-2.0, 2.0),
        "act_production_factor": (-0.3, 0.3),
        "act_price_factor": (-0.3, 0.3),
    }

    _log_mode = True

    @staticmethod
    def _make_obs(y_prev, yd, p, avg_price):
        import math

        def lg(x):
            return math.log(x) if x > 0 else (float("-inf") if x == 0 else float("nan"))

        return {"firm_stock": lg(y_prev + 1e-8) - lg(yd + 1e-8),
                "price_delta": lg(p + 1e-8) - lg(avg_price + 1e-8)}
