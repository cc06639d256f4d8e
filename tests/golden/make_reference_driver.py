"""Runs the REAL reference Python code (/root/reference/rmabm: Cats / CatsLog `reset` and `step`,
`Simulation`, `QLearner`, `Dummy`, `Logger`) and records what it returns:

    python tests/golden/make_reference_driver.py      ->  tests/golden/reference_driver.npz
    python tests/golden/make_reference_driver.py --reference-tests
        runs the reference's OWN test-suite (/root/reference/tests, unmodified, 33 tests) over the same
        stand-ins: a check that they honour the protocol the reference expects (33 passed).

The reference's Python layer needs two packages this image does not have: `gymnasium` (only its
`Env` base class, `spaces.Box/Dict` and `register`) and `juliacall` (the bridge to ABCredit.jl).  Both
are replaced here by small stand-ins.  The juliacall stand-in exposes `Main.ABCredit.<function>` for
the 21 functions `cats.py` calls; each one forwards to the matching phase of the in-repo CPU oracle
(`oracle/cats_oracle.c`, `ph_*`), and the model / firm / aggregate "Julia structs" are attribute
proxies onto the oracle's arrays.

So what these vectors pin is everything the reference does ITSELF on the path -- and nothing more:
  * the order of the 26 ABCredit calls of `Cats.step` and where the Python pieces sit between them
    (cats.py:355-419): a reward taken before tracking and bankruptcy, `agg.timestep += 1` from Python;
  * a6  `_implement_action` (cats.py:434-464, 673-704), incl. "dummy" entries, the alpha floor, both
    price-change modes, the log variants;
  * a18 `_get_reward` profits / rms (cats.py:493-556);  a21 `_get_terminated`;  a26 `_get_obs`
    (cats.py:466-481, 655-671);  a27 `_get_info` levels 1 and 2 (cats.py:558-626);
  * burn-in, `self.t` vs `agg.timestep`, truncation at T (cats.py:283-315, 423-432);
  * `Simulation.step(train=True)` with two `QLearner`s (numpy global RNG seeded by `reset`), and the
    `.npy` files `Logger` writes (simulation.py:93-146, qlearners.py:63-166, logger.py:42-202).
The arithmetic INSIDE the ABCredit calls is the oracle's (ABCredit.jl is not available here), so these
vectors do not pin that part: "parity unpinned" still holds for a1-a5, a7-a17, a19, a20, a22-a24.

This script only runs in the build container (it reads /root/reference).  Tests read the .npz.
"""
from __future__ import annotations

import ctypes as C
import sys
import tempfile
import types
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

REFERENCE = Path("/root/reference")


# ---------------------------------------------------------------------------------------------
# stand-in for gymnasium: just what cats.py / environments/__init__.py touch
def install_fake_gymnasium():
    gym = types.ModuleType("gymnasium")

    class Env:
        metadata: dict = {}

        def reset(self, seed=None, options=None):
            return None

        @property
        def unwrapped(self):
            return self

        def get_wrapper_attr(self, name):
            return getattr(self, name)

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape = low, high, shape or (1,)

        def sample(self):
            return np.random.uniform(self.low, self.high, size=self.shape).astype(np.float32)

    class Dict_(dict):
        shape = None          # as in gymnasium: a Dict space has no shape (heuristic.py:30 relies on it)

        def __init__(self, spaces_):
            super().__init__(spaces_)
            self.spaces = dict(spaces_)

        def sample(self):
            return {k: v.sample() for k, v in self.spaces.items()}

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box, spaces.Dict = Box, Dict_
    core = types.ModuleType("gymnasium.core")
    core.ObsType = core.ActType = core.RenderFrame = object
    gym.Env, gym.spaces, gym.core = Env, spaces, core
    registry = {}
    gym.register = lambda id, entry_point, **kw: registry.__setitem__(id, entry_point)

    def make(id, **kwargs):     # no wrappers: the tests only go through .unwrapped / get_wrapper_attr
        import importlib
        mod, cls = registry[id].split(":")
        return getattr(importlib.import_module(mod), cls)(**kwargs)
    gym.make = make
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.core": core})


# ---------------------------------------------------------------------------------------------
# stand-in for juliacall: ABCredit functions forwarded to the oracle's phases
class FirmRef:
    """A 'Julia firm struct': attribute reads / writes go to slot `i` of the oracle's arrays."""
    __slots__ = ("_m", "_kind", "_i")

    def __init__(self, model, kind, i):
        object.__setattr__(self, "_m", model)
        object.__setattr__(self, "_kind", kind)
        object.__setattr__(self, "_i", i)

    def __getattr__(self, name):
        m, kind, i = self._m, self._kind, self._i
        if name == "firm_id":
            return i + 1 if kind == "c" else m.F + i + 1
        if name == "Leff":
            return int(m.arr_i32(kind + "_Leff")[i])
        return float(m.arr_f64(f"{kind}_{name}")[i])

    def __setattr__(self, name, value):
        self._m.arr_f64(f"{self._kind}_{name}")[self._i] = float(value)

    def __eq__(self, other):
        return isinstance(other, FirmRef) and (self._m, self._kind, self._i) == (other._m, other._kind, other._i)


class WorkerRef:
    __slots__ = ("_m", "_i")

    def __init__(self, model, i):
        self._m, self._i = model, i

    def __eq__(self, other):
        return isinstance(other, WorkerRef) and (self._m, self._i) == (other._m, other._i)


class ScalRef:
    """agg / bank / gov: named scalars of the oracle's `scal` block."""

    def __init__(self, model, names):
        object.__setattr__(self, "_m", model)
        object.__setattr__(self, "_names", names)

    def __getattr__(self, name):
        v = float(self._m.arr_f64("scal")[orc.SCAL_NAMES.index(self._names[name])])
        return int(v) if name == "timestep" else v

    def __setattr__(self, name, value):
        self._m.arr_f64("scal")[orc.SCAL_NAMES.index(self._names[name])] = float(value)


class FakeModel:
    def __init__(self, W, F, N, params, seed):
        self.W, self.F, self.N = W, F, N
        self.econ = orc.OracleEconomy(W, F, N, params={k: params[k] for k in orc.PARAM_KEYS}, seed=seed, economy=0)
        self._views = {}
        self.params = dict(params)
        self.params["init_price"] = self.econ.scal("init_price")
        self.workers = [WorkerRef(self, i) for i in range(W)]
        self.consumption_firms = [FirmRef(self, "c", i) for i in range(F)]
        self.capital_firms = [FirmRef(self, "k", i) for i in range(N)]
        self.agg = ScalRef(self, {k: k for k in ("timestep", "price", "Y_real", "wb", "Un", "bankruptcy_rate",
                                                  "Y_nominal_tot", "gdp_deflator", "inflationRate",
                                                  "consumption", "totK")})
        self.bank = ScalRef(self, {"dividendsB": "dividendsB", "profitsB": "profitsB"})
        self.gov = ScalRef(self, {"GB": "GB", "deficitGDP": "deficitGDP", "bonds": "gov_bonds"})

    def arr_f64(self, name):
        v = self._views.get(name)
        if v is None:
            v = self._views[name] = self.econ.f64(name)
        return v

    def arr_i32(self, name):
        v = self._views.get(name)
        if v is None:
            v = self._views[name] = self.econ.i32(name)
        return v


CALLS: list[str] = []          # every ABCredit call, in order: the test of the sequencing


def install_fake_juliacall():
    jc = types.ModuleType("juliacall")

    class JuliaError(Exception):
        pass

    L = orc.lib()
    state = {"seed": 0}

    def ph(name):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = None, [C.c_void_p]
        return fn

    def is_k(firms):
        return firms[0]._kind == "k"

    def by_kind(c_name, k_name):
        fc, fk = ph(c_name), ph(k_name)

        def call(firms, model=None):
            m = firms[0]._m
            (fk if is_k(firms) else fc)(m.econ._h)
        return call

    def on_model(name, model_arg=-1):
        f = ph(name)

        def call(*args):
            m = args[model_arg]
            m = m if isinstance(m, FakeModel) else m._m
            f(m.econ._h)
        return call

    class ABCredit:
        pass

    def rec(name, fn):
        def wrapped(*a):
            CALLS.append(name)
            return fn(*a)
        setattr(ABCredit, name, staticmethod(wrapped))

    def initialise_model(W, F, N, params):
        seed = state["seed"]
        return FakeModel(W, F, N, params, 0 if seed is None else int(seed))
    ABCredit.initialise_model = staticmethod(initialise_model)
    rec("reset_gov_and_banking_variables_b", on_model("ph_reset_gov_bank"))
    rec("set_bond_interest_rate_b", on_model("ph_bond_rate"))
    pq = by_kind("ph_price_quantity_c", "ph_price_quantity_k")
    rec("firms_decide_price_quantity_b", pq)
    rec("firms_decide_investment_b", on_model("ph_investment"))
    rec("firms_decide_labour_b", by_kind("ph_labour_c", "ph_labour_k"))
    rec("firms_get_credit_b", by_kind("ph_credit_c", "ph_credit_k"))
    rec("firms_fire_workers_b", by_kind("ph_fire_c", "ph_fire_k"))
    rec("workers_search_job_b", on_model("ph_job_search"))
    rec("firms_produce_b", by_kind("ph_produce_c", "ph_produce_k"))
    rec("capital_goods_market_b", on_model("ph_capital_market"))
    rec("gov_adjusts_subsidy_b", on_model("ph_gov_subsidy"))
    rec("workers_get_paid_b", on_model("ph_get_paid"))
    rec("households_find_cons_budget_b", on_model("ph_cons_budget"))
    rec("consumption_goods_market_b", on_model("ph_goods_market"))
    rec("firms_accounting_b", by_kind("ph_accounting_c", "ph_accounting_k"))
    rec("update_tracking_variables_b", on_model("ph_tracking"))
    rec("firms_go_bankrupt_b", on_model("ph_bankrupt"))
    rec("bank_accounting_b", on_model("ph_bank_accounting"))
    rec("gov_accounting_b", on_model("ph_gov_accounting"))
    rec("adjust_wages_b", on_model("ph_adjust_wages"))

    class Random:
        @staticmethod
        def seed_b(seed):
            state["seed"] = seed

    class Main:
        pass
    Main.ABCredit, Main.Random = ABCredit, Random
    Main.seval = staticmethod(lambda s: None)
    Main.cat = staticmethod(lambda *vs, dims=1: [x for v in vs for x in v])
    jc.Main, jc.JuliaError, jc.VectorValue = Main, JuliaError, list
    jc.convert = lambda T=None, x=None: dict(x)
    sys.modules["juliacall"] = jc


# ---------------------------------------------------------------------------------------------
INFO_SCALARS = ["Y_real", "wb", "Un", "bankruptcy_rate", "avg_price", "Y_nominal_tot", "gdp_deflator",
                "inflationRate", "consumption", "totK", "dividendsB", "profitsB", "GB", "deficitGDP", "bonds"]
INFO_FIRMS = ["sales", "production", "debt", "employment", "capital", "equity", "investment", "liquidity"]


def flatten_info(info: dict, n_agents: int) -> np.ndarray:
    """One row per step: the 15 scalars (NaN where the level does not have them), then for each of
    the 8 per-firm quantities the agents' values followed by (mean, std) of the others."""
    row = [float(info.get(k, np.nan)) for k in INFO_SCALARS]
    for q in INFO_FIRMS:
        ag = info.get("agents_" + q)
        ot = info.get("others_" + q)
        row += [float(x) for x in ag] if ag is not None else [np.nan] * n_agents
        row += [float(ot[0]), float(ot[1])] if ot is not None else [np.nan, np.nan]
    return np.array(row, dtype=np.float64)


def obs_array(obs) -> np.ndarray:
    return np.array([[o["firm_stock"], o["price_delta"]] for o in obs], dtype=np.float64).reshape(-1, 2)


SCENARIOS = [
    # name, class, ctor kwargs, seed, action script (cycled); None = ("dummy", "dummy")
    ("cats_profits", "Cats", dict(T=12, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=2), 11,
     [[(1.1, 0.9), None], [(0.95, 1.05), (1.2, 1.0)], [None, None], [(1.0, 1.0), (0.7, 1.3)]]),
    ("cats_avg_rms", "Cats", dict(T=8, W=300, F=30, N=6, t_burnin=20, n_agents=3, info_level=1,
                                   reward_type="rms", price_change="avg_price"), 5,
     [[(1.05, 0.98), (0.9, 1.1), None], [(1.3, 0.7), None, (1.0, 1.02)]]),
    ("cats_floor", "Cats", dict(T=6, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=2), 3,
     [[(0.0, 0.0), (1.0, 1.0)], [(0.0, 0.0), None]]),      # reference tests/test_env.py:157-178: De -> alpha, P -> 0
    ("catslog_profits", "CatsLog", dict(T=10, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=2), 7,
     [[(0.1, -0.1), None], [(-0.05, 0.02), (0.3, -0.3)], [(0.0, 0.0), (-0.3, 0.3)]]),
    ("catslog_avg_rms", "CatsLog", dict(T=6, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=1,
                                         reward_type="rms", price_change="avg_price"), 2,
     [[(0.2, 0.05), (-0.1, -0.05)], [None, (0.05, 0.0)]]),
    ("default_dummy", "Cats", dict(T=5, n_agents=1, info_level=2), 0,      # examples/heuristic_run.py: one Dummy firm
     [[None]]),
]


def run_env_scenarios(envs, out):
    for name, cls, kw, seed, script in SCENARIOS:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = getattr(envs, cls)(**kw)
            n = kw["n_agents"]
            obs, info = env.reset(seed=seed)
            out[f"{name}/reset_obs"] = obs_array(obs)
            out[f"{name}/reset_info"] = flatten_info(info, n)
            del CALLS[:]
            O, R, TT, I, A, M = [], [], [], [], [], []
            for t in range(kw["T"]):
                acts = [("dummy", "dummy") if a is None else a for a in script[t % len(script)]]
                obs, rew, term, trunc, info = env.step(acts)
                O.append(obs_array(obs)); R.append(np.array(rew, dtype=np.float64))
                TT.append([int(term), int(trunc), env.t, int(env.model.agg.timestep)])
                I.append(flatten_info(info, n))
                A.append([[1.0, 1.0] if a[0] == "dummy" else [float(a[0]), float(a[1])] for a in acts])
                M.append([0 if a[0] == "dummy" else 1 for a in acts])
            out[f"{name}/obs"], out[f"{name}/reward"] = np.array(O), np.array(R)
            out[f"{name}/flags"], out[f"{name}/info"] = np.array(TT, dtype=np.int64), np.array(I)
            out[f"{name}/actions"], out[f"{name}/mask"] = np.array(A, dtype=np.float64), np.array(M, dtype=np.uint8)
            # final state of the C firms the Python layer can see, as a cross-check
            m = env.model
            out[f"{name}/final_De"] = m.arr_f64("c_De").copy()
            out[f"{name}/final_P"] = m.arr_f64("c_P").copy()
            out[f"{name}/final_A"] = m.arr_f64("c_A").copy()
        if name == "cats_profits":
            out["call_order"] = np.array(CALLS[:len(CALLS) // kw["T"]])
        print(name, "steps", kw["T"], "reward[0]", out[f"{name}/reward"][0])


def run_simulation_scenario(rmabm, out):
    """simulation.py + qlearners.py + logger.py over the env: 2 Q-learners trained for one episode."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # linear mode: no libm call anywhere on the path, so the whole episode is reproducible to the bit
        env = rmabm.environments.Cats(T=60, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=1)
        agents = [rmabm.agents.QLearner(i, env, n_bins=(5, 5, 3, 3), alpha=0.2, gamma=0.9, epsilon_zero=0.3)
                  for i in env.agents_ids]
        sim = rmabm.Simulation(env, agents, n_episodes=1)
        with tempfile.TemporaryDirectory() as d:
            sim.init_logger("ref", d, use_timestamp=False)
            sim.reset(seed=21)
            while not sim._terminated and not sim._truncated:
                sim.step(train=True)
            sim.logger.log_obs(sim._obs_history, 1)
            sim.logger.log_act(sim._act_history, 1)
            sim.logger.log_array(np.array(sim._reward_history), "ep1_rewards")
            sim.logger.log_info(sim._info_history, 1)
            for f in sorted(Path(d).rglob("*.npy")):
                out["sim/file/" + f.name] = np.load(f, allow_pickle=False)
        for i, ag in enumerate(agents):
            out[f"sim/Q{i}"] = ag.Q.copy()
    print("simulation: files", sorted(k for k in out if k.startswith("sim/file/")))


def main():
    install_fake_gymnasium()
    install_fake_juliacall()
    sys.path.insert(0, str(REFERENCE))
    if "--reference-tests" in sys.argv:
        import pytest
        raise SystemExit(pytest.main([str(REFERENCE / "tests"), "-p", "no:cacheprovider", "-q",
                                      "--rootdir", str(REFERENCE)]))
    import rmabm                     # the real reference package
    assert Path(rmabm.__file__).resolve().is_relative_to(REFERENCE)
    out: dict[str, np.ndarray] = {}
    run_env_scenarios(rmabm.environments, out)
    run_simulation_scenario(rmabm, out)
    out["info_scalars"] = np.array(INFO_SCALARS)
    out["info_firms"] = np.array(INFO_FIRMS)
    out["recorded_with"] = np.array(["python " + sys.version.split()[0], "numpy " + np.__version__])
    np.savez_compressed(HERE / "reference_driver.npz", **out)
    print("written", HERE / "reference_driver.npz", (HERE / "reference_driver.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
