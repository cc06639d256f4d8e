"""Model parameters: defaults and loaders.

Mirrors `Cats._default_parameters` and `Cats._load_parameters`
(/root/reference/rmabm/environments/cats.py:69-104, 175-219): None / dict / .json / .csv with
default-fill, in-place fill of a caller's dict (pinned by the reference's tests/test_env.py:124-135),
ValueError for a bad suffix or CSV length.
"""
from __future__ import annotations

import csv
import json
from pathlib import Path
from typing import Any, Dict

DEFAULT_PARAMETERS: Dict[str, Any] = {
    "z_c": 5,  # no. of applications in consumption good market
    "z_k": 2,  # no. of applications in capital good market
    "z_e": 5,  # number of job applications
    "xi": 0.96,  # memory parameter human wealth
    "chi": 0.05,  # fraction of wealth devoted to consumption
    "q_adj": 0.9,  # quantity adjustment parameter
    "p_adj": 0.1,  # price adjustment parameter
    "mu": 1.2,  # bank's gross mark-up
    "eta": 0.03,  # capital depreciation
    "Iprob": 0.25,  # probability of investing
    "phi": 0.002,  # bank's leverage parameter
    "theta": 0.05,  # rate of debt reimbursement
    "delta": 0.5,  # memory parameter in the capital utilization rate
    "alpha": 0.66667,  # labour productivity
    "k": 0.33333,  # capital productivity
    "div": 0.2,  # share of dividends
    "barX": 0.85,  # desired capital utilization
    "inventory_depreciation": 0.3,  # rate at which capital firms' inventories depreciate
    "b1": -15,  # parameters for risk evaluation by banks
    "b2": 13,
    "b_k1": -5,
    "b_k2": 5,
    "interest_rate": 0.01,
    "subsidy": 0,
    "maastricht": 0.0,
    "target_deficit": 0.01,
    "tax_rate": 0.0,
    "wage_update_up": 0.1,
    "wage_update_down": 0.01,
    "u_target": 0.1,
    "wb": 1.0,  # initial wage rate
    "tax_rate_d": 0.0,  # taxes on dividends
    "r_f": 0.01,  # general refinancing rate
    "E_threshold_scale": 0,
}
PARAM_KEYS = list(DEFAULT_PARAMETERS.keys())
assert len(PARAM_KEYS) == 34


def load_parameters(params: Dict[str, Any] | str | Path | None) -> Dict[str, Any]:
    """Return the full 34-key parameter dict for `params` given as None, dict, .json or .csv path."""
    if params is None:
        return dict(DEFAULT_PARAMETERS)
    if isinstance(params, dict):
        for key, value in DEFAULT_PARAMETERS.items():  # fills the CALLER's dict, as the reference does
            if key not in params:
                params[key] = value
        return params
    if isinstance(params, str):
        params = Path(params)
    if isinstance(params, Path):
        if params.suffix == ".json":
            with open(params, "r") as f:
                loaded = json.load(f)
            for key, value in DEFAULT_PARAMETERS.items():
                if key not in loaded:
                    loaded[key] = value
            return loaded
        if params.suffix == ".csv":
            with open(params, "r") as f:
                values = [float(row[0]) for row in csv.reader(f)]
            values[:3] = [int(x) for x in values[:3]]  # z_c, z_k, z_e are integers
            if len(values) != len(PARAM_KEYS):
                raise ValueError(f"Invalid number of parameters. Expected {len(PARAM_KEYS)}, got {len(values)}")
            return dict(zip(PARAM_KEYS, values))
        raise ValueError(f"Invalid file type. Must be .json or .csv. Got {params.suffix}")
    raise ValueError(f"Invalid params argument of type {type(params).__name__}")


def params_vector(params: Dict[str, Any]) -> list[float]:
    """34 doubles in the C-ABI order (include/cats_b200.h CATS_N_PARAMS)."""
    return [float(params[k]) for k in PARAM_KEYS]


def validate(params: Dict[str, Any], F: int, N: int) -> None:
    for key in ("z_c", "z_k", "z_e"):
        z = int(params[key])
        if not 1 <= z <= 8:
            raise ValueError(f"{key} must be in [1, 8] (got {z})")
    if not 0 < float(params["theta"]) < 1:
        raise ValueError("theta must be in (0, 1)")
    for key in ("alpha", "k", "wb", "barX", "Iprob"):
        if not float(params[key]) > 0:
            raise ValueError(f"{key} must be > 0")
