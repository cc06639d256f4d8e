"""Parity against vectors recorded from the REAL reference Python code
(tests/golden/reference_driver.npz, written by tests/golden/make_reference_driver.py: the reference's
`Cats.step` / `CatsLog.step`, `Simulation`, `QLearner`, `Logger` run unmodified over a juliacall
stand-in whose ABCredit functions are the oracle's phases).

What is pinned here is the part of the path the reference implements itself: the call order of
cats.py:355-419, the action override (a6), reward (a18), termination (a21), observations (a26),
info (a27), burn-in / time bookkeeping, and the agent / logger layer above.  CPU tests check the
oracle's `oracle_step` (one C function for the whole period) and the host-side mirror; the GPU tests
check the `Cats` / `CatsLog` adapters over the CUDA engine, through the C-ABI.

Tolerances: linear mode is bit-exact.  Log mode (`CatsLog`) observations and actions go through
`math.log` / `math.exp` in the reference and through the model's own correctly-specified
`det_log` / `det_exp` in the engine and the oracle (docs/MODEL.md 4.2): LOG_RTOL below, well inside the
1e-9 relative band of the brief.  The market-share reward (`reward_type="rms"`) divides by
`sum([...])` over Python floats (cats.py:548-549); CPython >= 3.12 -- the interpreter that recorded these
vectors -- adds them with compensated summation, and so do the oracle and the kernel: bit-exact too.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = Path(__file__).resolve().parent / "golden" / "reference_driver.npz"
LOG_RTOL, LOG_ATOL = 1e-12, 1e-13

# name -> (class name, ctor kwargs, seed); the action scripts are in the file
SCENARIOS = {
    "cats_profits": ("Cats", dict(T=12, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=2), 11),
    "cats_avg_rms": ("Cats", dict(T=8, W=300, F=30, N=6, t_burnin=20, n_agents=3, info_level=1,
                                   reward_type="rms", price_change="avg_price"), 5),
    "cats_floor": ("Cats", dict(T=6, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=2), 3),
    "catslog_profits": ("CatsLog", dict(T=10, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=2), 7),
    "catslog_avg_rms": ("CatsLog", dict(T=6, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=1,
                                         reward_type="rms", price_change="avg_price"), 2),
    "default_dummy": ("Cats", dict(T=5, n_agents=1, info_level=2), 0),
}

ABCREDIT_CALL_ORDER = [      # cats.py:355-410
    "reset_gov_and_banking_variables_b", "set_bond_interest_rate_b",
    "firms_decide_price_quantity_b", "firms_decide_price_quantity_b", "firms_decide_investment_b",
    "firms_decide_labour_b", "firms_decide_labour_b", "firms_get_credit_b", "firms_get_credit_b",
    "firms_fire_workers_b", "firms_fire_workers_b", "workers_search_job_b",
    "firms_produce_b", "firms_produce_b", "capital_goods_market_b", "gov_adjusts_subsidy_b",
    "workers_get_paid_b", "households_find_cons_budget_b", "consumption_goods_market_b",
    "firms_accounting_b", "firms_accounting_b", "update_tracking_variables_b", "firms_go_bankrupt_b",
    "bank_accounting_b", "gov_accounting_b", "adjust_wages_b",
]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=False)


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def check(a, b, log_mode, what, rms=False):
    if log_mode:
        np.testing.assert_allclose(a, b, rtol=LOG_RTOL, atol=LOG_ATOL, equal_nan=True, err_msg=what)
    else:
        assert same_bits(a, b), f"{what}: {np.asarray(a).ravel()[:6]} vs {np.asarray(b).ravel()[:6]}"


def flags_of(cls, kw):
    f = orc.FLAG_LOG if cls == "CatsLog" else 0
    if kw.get("price_change") == "avg_price":
        f |= orc.FLAG_AVG_PRICE
    if kw.get("reward_type") == "rms":
        f |= orc.FLAG_REWARD_RMS
    return f


def test_reference_calls_abcredit_in_the_documented_order(gold):
    """The 26 calls of one `Cats.step`, as the reference made them."""
    assert [str(s) for s in gold["call_order"]] == ABCREDIT_CALL_ORDER


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_oracle_step_matches_reference_python_layer(gold, name):
    """`oracle_step` (sequencing + a6 + a18 + a26 in C) == the reference's `step` driving the phases
    one by one from Python."""
    cls, kw, seed = SCENARIOS[name]
    log_mode = cls == "CatsLog"
    W, F, N = kw.get("W", 1000), kw.get("F", 100), kw.get("N", 20)
    n, fl = kw["n_agents"], flags_of(cls, kw)
    e = orc.OracleEconomy(W, F, N, seed=seed, economy=0, bankruptcy_reward=-100.0)
    for _ in range(kw.get("t_burnin", 300)):
        e.step()
    check(e.obs(n, fl), gold[f"{name}/reset_obs"], log_mode, "reset obs")
    sc = gold["info_scalars"].tolist()
    for t in range(kw["T"]):
        obs, rew = e.step(actions=gold[f"{name}/actions"][t], mask=gold[f"{name}/mask"][t], flags=fl)
        check(obs, gold[f"{name}/obs"][t], log_mode, f"obs step {t}")
        check(rew, gold[f"{name}/reward"][t], log_mode, f"reward step {t}", rms=kw.get("reward_type") == "rms")
        term, trunc, env_t, agg_t = gold[f"{name}/flags"][t]
        assert int(e.scal("terminated")) == term
        assert int(e.scal("timestep")) == agg_t == kw.get("t_burnin", 300) + t + 1 and env_t == t + 1
        assert trunc == int(t + 1 >= kw["T"])
        info = gold[f"{name}/info"][t]
        for key, okey in (("Y_real", "Y_real"), ("wb", "wb"), ("Un", "Un"), ("bankruptcy_rate", "bankruptcy_rate"),
                          ("avg_price", "price")):
            check(info[sc.index(key)], e.scal(okey), log_mode, key)
    # log mode: the override itself goes through exp/log, so the state may differ in the last bits
    check(e.f64("c_De"), gold[f"{name}/final_De"], log_mode, "final De")
    check(e.f64("c_P"), gold[f"{name}/final_P"], log_mode, "final P")
    check(e.f64("c_A"), gold[f"{name}/final_A"], log_mode, "final A")


class OracleEnv:
    """Test-only environment with the reference's `reset` / `step` shape over the CPU oracle, so the
    host-side agent / simulation / logger mirror can be checked without a GPU."""

    def __init__(self, log_mode, T, W, F, N, t_burnin, n_agents, bounds):
        self.log_mode, self.T, self.W, self.F, self.N = log_mode, T, W, F, N
        self.t_burnin, self.n_agents, self.gym_spaces_bounds = t_burnin, n_agents, bounds
        self.agents_ids = list(range(1, n_agents + 1))
        self.fl = orc.FLAG_LOG if log_mode else 0
        self.t = 0

    def _obs(self):
        return [{"firm_stock": float(o[0]), "price_delta": float(o[1])} for o in self.e.obs(self.n_agents, self.fl)]

    def _info(self):
        Q = self.e.f64("c_Q")
        n = self.n_agents
        return {"Y_real": self.e.scal("Y_real"), "wb": self.e.scal("wb"), "Un": self.e.scal("Un"),
                "bankruptcy_rate": self.e.scal("bankruptcy_rate"), "avg_price": self.e.scal("price"),
                "agents_sales": [float(x) for x in Q[:n]], "others_sales": (np.mean(Q[n:]), np.std(Q[n:]))}

    def reset(self, seed=None, options=None):
        import random
        random.seed(seed); np.random.seed(seed)
        self.e = orc.OracleEconomy(self.W, self.F, self.N, seed=0 if seed is None else seed, economy=0)
        for _ in range(self.t_burnin):
            self.e.step()
        self.t = 0
        return self._obs(), self._info()

    def step(self, actions):
        a = [[1.0, 1.0] if "dummy" in x else [float(x[0]), float(x[1])] for x in actions]
        m = [0 if "dummy" in x else 1 for x in actions]
        obs, rew = self.e.step(actions=a, mask=m, flags=self.fl)
        self.t += 1
        return (self._obs(), [float(r) for r in rew], bool(self.e.scal("terminated")), self.t >= self.T, self._info())


LIN_BOUNDS = {"obs_firm_stock": (-5.0, 5.0), "obs_price_delta": (-2.0, 8.0),
              "act_production_factor": (0.7, 1.3), "act_price_factor": (0.7, 1.3)}


def run_our_simulation(env, tmp_path):
    from r_mabm_b200.agents import QLearner
    from r_mabm_b200.simulation import Simulation
    agents = [QLearner(i, env, n_bins=(5, 5, 3, 3), alpha=0.2, gamma=0.9, epsilon_zero=0.3) for i in env.agents_ids]
    sim = Simulation(env, agents, n_episodes=1)
    sim.init_logger("ours", tmp_path, use_timestamp=False)
    sim.reset(seed=21)
    while not sim._terminated and not sim._truncated:
        sim.step(train=True)
    sim.logger.log_obs(sim._obs_history, 1)
    sim.logger.log_act(sim._act_history, 1)
    sim.logger.log_array(np.array(sim._reward_history), "ep1_rewards")
    sim.logger.log_info(sim._info_history, 1)
    files = {f.name: np.load(f, allow_pickle=False) for f in sorted(Path(tmp_path).rglob("*.npy"))}
    return files, [a.Q for a in agents]


def compare_simulation(files, Qs, gold):
    ref_files = {k[len("sim/file/"):]: gold[k] for k in gold.files if k.startswith("sim/file/")}
    assert sorted(files) == sorted(ref_files)
    # linear mode, profits reward: every logged number and both Q tables are identical to the bit
    for name, ref in ref_files.items():
        assert files[name].shape == ref.shape and files[name].dtype == ref.dtype, name
        assert same_bits(files[name], ref), name
    for i, Q in enumerate(Qs):
        assert same_bits(Q, gold[f"sim/Q{i}"])
        assert np.count_nonzero(gold[f"sim/Q{i}"]) > 5      # the episode did train


def test_simulation_qlearner_logger_match_reference_files(gold, tmp_path):
    """Our Simulation + QLearner + Logger over an oracle-backed env write the same .npy files
    (names, shapes, dtypes, values) as the reference's classes did, and end with the same Q tables."""
    env = OracleEnv(False, T=60, W=300, F=30, N=6, t_burnin=20, n_agents=2, bounds=dict(LIN_BOUNDS))
    files, Qs = run_our_simulation(env, tmp_path)
    compare_simulation(files, Qs, gold)


# ---------------------------------------------------------------------------------------------
def unflatten_info(info, n_agents, gold_scalars, gold_firms):
    """What make_reference_driver.flatten_info would produce for `info`."""
    row = [float(info.get(k, np.nan)) for k in gold_scalars]
    for q in gold_firms:
        ag, ot = info.get("agents_" + q), info.get("others_" + q)
        row += [float(x) for x in ag] if ag is not None else [np.nan] * n_agents
        row += [float(ot[0]), float(ot[1])] if ot is not None else [np.nan, np.nan]
    return np.array(row, dtype=np.float64)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_adapter_matches_reference_step_outputs(gold, name):
    """`Cats` / `CatsLog` over the CUDA engine return what the reference's `reset` / `step` returned:
    observations, rewards, terminated, truncated, and every info key of the level."""
    import warnings
    from r_mabm_b200 import environments as envs
    cls, kw, seed = SCENARIOS[name]
    log_mode = cls == "CatsLog"
    n = kw["n_agents"]
    sc, fm = gold["info_scalars"].tolist(), gold["info_firms"].tolist()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = getattr(envs, cls)(**kw)
        obs, info = env.reset(seed=seed)
        check(np.array([[o["firm_stock"], o["price_delta"]] for o in obs]).reshape(-1, 2),
              gold[f"{name}/reset_obs"], log_mode, "reset obs")
        check(unflatten_info(info, n, sc, fm), gold[f"{name}/reset_info"], log_mode, "reset info")
        for t in range(kw["T"]):
            acts = [("dummy", "dummy") if not m else (float(a[0]), float(a[1]))
                    for a, m in zip(gold[f"{name}/actions"][t], gold[f"{name}/mask"][t])]
            obs, rew, term, trunc, info = env.step(acts)
            assert isinstance(term, bool) and isinstance(trunc, bool) and isinstance(rew, list)
            check(np.array([[o["firm_stock"], o["price_delta"]] for o in obs]).reshape(-1, 2),
                  gold[f"{name}/obs"][t], log_mode, f"obs step {t}")
            check(np.array(rew), gold[f"{name}/reward"][t], log_mode, f"reward step {t}",
                  rms=kw.get("reward_type") == "rms")
            g_term, g_trunc, g_t, _ = gold[f"{name}/flags"][t]
            assert (int(term), int(trunc), env.t) == (g_term, g_trunc, g_t)
            check(unflatten_info(info, n, sc, fm), gold[f"{name}/info"][t], log_mode, f"info step {t}")
        check(env._field("c_De"), gold[f"{name}/final_De"], log_mode, "final De")
        check(env._field("c_P"), gold[f"{name}/final_P"], log_mode, "final P")
        check(env._field("c_A"), gold[f"{name}/final_A"], log_mode, "final A")


@pytest.mark.gpu
def test_simulation_over_the_engine_matches_reference_files(gold, tmp_path):
    """The reference's training loop, switched to this framework: Simulation + 2 QLearners + Logger over
    `Cats` on the GPU reproduce the reference's logged episode and Q tables."""
    import warnings
    from r_mabm_b200.environments import Cats
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = Cats(T=60, W=300, F=30, N=6, t_burnin=20, n_agents=2, info_level=1)
        files, Qs = run_our_simulation(env, tmp_path)
    compare_simulation(files, Qs, gold)
