"""CPU: host-side mirror of the reference interface -- parameter loaders, agents, logger, spaces,
Simulation bookkeeping.  Known answers are the ones the reference's own tests pin
(/root/reference/tests/test_agents.py:38-108, test_env.py:50-155, test_logger.py:23-240)."""
from pathlib import Path

import numpy as np
import pytest

from r_mabm_b200.agents import Dummy, QLearner
from r_mabm_b200.environments import spaces
from r_mabm_b200.logger import Logger
from r_mabm_b200.params import DEFAULT_PARAMETERS, PARAM_KEYS, load_parameters, params_vector, validate
from r_mabm_b200.rollout import shard_range
from r_mabm_b200.simulation import Simulation

RES = Path(__file__).resolve().parent / "resources"


class FakeEnv:
    """Stands in for Cats on a box without a GPU: same duck type, trivial dynamics."""
    gym_spaces_bounds = {"obs_firm_stock": (-5.0, 5.0), "obs_price_delta": (-2.0, 8.0),
                         "act_production_factor": (0.7, 1.3), "act_price_factor": (0.7, 1.3)}

    def __init__(self, n_agents=2, T=10):
        self.n_agents, self.T, self.t = n_agents, T, 0
        self.action_space = spaces.Dict({"production_factor": spaces.Box(0.7, 1.3), "price_factor": spaces.Box(0.7, 1.3)})
        self.resets = 0

    def _obs(self):
        return [{"firm_stock": 0.1 * self.t, "price_delta": -0.2 * self.t} for _ in range(self.n_agents)]

    def reset(self, seed=None, options=None):
        self.t = 0; self.resets += 1
        return self._obs(), {"Y_real": 1.0}

    def step(self, actions):
        assert len(actions) == self.n_agents
        self.t += 1
        return self._obs(), [1.0 + i for i in range(self.n_agents)], False, self.t >= self.T, {"Y_real": float(self.t), "Un": 0.1}


def test_default_parameters_and_vector_order():
    assert len(PARAM_KEYS) == 34 and PARAM_KEYS[:3] == ["z_c", "z_k", "z_e"] and PARAM_KEYS[-1] == "E_threshold_scale"
    p = load_parameters(None)
    assert p["alpha"] == 0.66667 and p["b1"] == -15 and p["wb"] == 1.0
    v = params_vector(p)
    assert v[0] == 5.0 and v[13] == 0.66667 and len(v) == 34
    validate(p, 100, 20)
    with pytest.raises(ValueError):
        validate({**p, "z_c": 9}, 100, 20)


def test_partial_dict_is_filled_in_place():
    mine = {"z_c": 4, "chi": 0.07}
    out = load_parameters(mine)
    assert out is mine and mine["z_c"] == 4 and mine["z_k"] == 2 and len(mine) == 34


def test_json_and_csv_loaders_and_errors(tmp_path):
    j = load_parameters(RES / "test_parameters.json")
    assert j["z_c"] == 4 and j["z_k"] == DEFAULT_PARAMETERS["z_k"] and len(j) == 34
    c = load_parameters(str(RES / "test_parameters.csv"))
    assert c["z_c"] == 4 and isinstance(c["z_c"], int) and isinstance(c["z_e"], int) and c["xi"] == 0.96
    bad = tmp_path / "p.txt"; bad.write_text("1\n")
    with pytest.raises(ValueError):
        load_parameters(bad)
    short = tmp_path / "p.csv"; short.write_text("1\n2\n3\n")
    with pytest.raises(ValueError):
        load_parameters(short)


def test_qlearner_known_answers():
    env = FakeEnv()
    np.random.seed(0)
    ag = QLearner(7, env, n_bins=(11, 11, 11, 11), alpha=0.1, gamma=0.99, epsilon_zero=0.0)
    assert ag.agent_id == 7 and ag.Q.shape == (11, 11, 11, 11)
    # nearest pole, clamping, first index on ties
    assert ag.bin_obs({"firm_stock": -5.0, "price_delta": -2.0}) == (0, 0)
    assert ag.bin_obs({"firm_stock": 5.0, "price_delta": 8.0}) == (10, 10)
    assert ag.bin_obs({"firm_stock": 0.0, "price_delta": 3.0}) == (5, 5)
    assert ag.bin_obs({"firm_stock": 100.0, "price_delta": -100.0}) == (10, 0)
    assert ag.bin_obs({"firm_stock": 0.49, "price_delta": 3.51}) == (5, 6)
    # greedy action on a marked table: index (a0, a1) -> lo + idx * (hi - lo) / 10
    ag.Q[5, 5, 10, 8] = 1.0
    act = ag.get_action({"firm_stock": 0.0, "price_delta": 3.0})
    assert act == pytest.approx((1.3, 0.7 + 8 * 0.06))
    # TD(0): Q += alpha (r + gamma max Q[s'] - Q); no bootstrap when terminated
    ag.Q[6, 6, 2, 2] = 2.0
    ag.train({"firm_stock": 0.0, "price_delta": 3.0}, 1.0, {"firm_stock": 1.0, "price_delta": 4.0}, False)
    assert ag.Q[5, 5, 10, 8] == pytest.approx(1.0 + 0.1 * (1.0 + 0.99 * 2.0 - 1.0))
    before = ag.Q[5, 5, 10, 8]
    ag.train({"firm_stock": 0.0, "price_delta": 3.0}, 3.0, None, True)
    assert ag.Q[5, 5, 10, 8] == pytest.approx(before + 0.1 * (3.0 - before))
    with pytest.raises(TypeError):
        ag.train({"firm_stock": 0.0, "price_delta": 3.0}, 3.0, None, False)
    with pytest.warns(UserWarning):
        QLearner(1, env, alpha=1.5)


def test_dummy_agent():
    d = Dummy(3, FakeEnv())
    assert d.get_action({"anything": 1}) == ("dummy", "dummy") and d.bin_obs() == (-1, -1)
    assert d.train(1, 2, 3) is None and Dummy(0).action_length == 2


def test_logger_layout(tmp_path):
    lg = Logger("run", tmp_path, use_timestamp=False)
    assert lg.path == tmp_path / "run" and lg.path.is_dir()
    assert Logger("run", tmp_path, use_timestamp=True).path.name.startswith("run_20")
    lg.log_array(np.zeros(0), "empty")
    assert not (lg.path / "empty.npy").exists()
    lg.log_array(np.arange(3), "a"); lg.log_array(np.arange(2), "a.npy")
    assert np.load(lg.path / "a.npy").tolist() == [0, 1, 2, 0, 1]
    infos = [{"Y_real": float(t), "others_sales": (1.0, 2.0)} for t in range(10)]
    lg.log_info(infos, 1)
    assert np.load(lg.path / "ep1_Y_real.npy").shape == (10,) and np.load(lg.path / "ep1_others_sales.npy").shape == (10, 2)
    lg.log_info([], 2); lg.log_info([{}], 2)
    obs = [[{"firm_stock": 1.0, "price_delta": 2.0}] * 3 for _ in range(10)]
    lg.log_obs(obs, 1)
    assert np.load(lg.path / "ep1_obs.npy").shape == (10, 3, 2)
    acts = [[(1.1, 0.9), ("dummy", "dummy")] for _ in range(10)]
    lg.log_act(acts, 1)
    a = np.load(lg.path / "ep1_act.npy")
    assert a.shape == (10, 2, 2) and (a[:, 1] == 0).all() and a[0, 0].tolist() == [1.1, 0.9]
    lg.log_act([[("dummy", "dummy")]] * 4, 3)
    assert not (lg.path / "ep3_act.npy").exists()
    with pytest.raises(TypeError):
        lg._log_gym([[1.0]], "x")
    stats = np.arange(2 * 5 * 16, dtype=np.float64).reshape(2, 5, 16)
    lg.log_batched_stats(stats, episode=4, keys=["Un"])
    assert np.load(lg.path / "ep4_Un.npy").shape == (5, 2)


def test_simulation_bookkeeping(tmp_path):
    env = FakeEnv(n_agents=2, T=10)
    with pytest.raises(ValueError):
        Simulation(env, [Dummy(0)])
    agents = [QLearner(1, env, epsilon_zero=1.0), Dummy(2, env)]
    sim = Simulation(env, agents, n_episodes=3)
    assert env.resets == 1 and sim.current_episode == 1
    sim.step(train=True)
    assert len(sim.obs_history) == 1 and len(sim.reward_history[0]) == 2 and np.any(agents[0].Q != 0)
    sim.init_logger("log", tmp_path, use_timestamp=False)
    sim.run_episode()
    assert sim.current_episode == 2 and env.resets == 2 and sim.obs_history == []
    assert np.load(tmp_path / "log" / "ep1_rewards.npy").shape == (10, 2)
    assert np.load(tmp_path / "log" / "ep1_Y_real.npy").shape == (10,)
    assert np.load(tmp_path / "log" / "ep1_act.npy").shape == (10, 2, 2)
    sim.train_agents()
    assert sim.current_episode == 4 and agents[0].epsilon == pytest.approx(max(0.01, 0.9 ** 4))
    sim.train_agents(n_episodes=1, update=lambda a, s: setattr(a, "epsilon", 0.5))
    assert agents[0].epsilon == 0.5


def test_spaces_stand_ins():
    d = spaces.Dict({"a": spaces.Box(low=-1.0, high=2.0), "b": spaces.Box(low=0.0, high=1.0)})
    assert len(d.keys()) == 2 and d["a"].low.item(0) == -1.0 and d["a"].high.item(0) == 2.0
    s = d.sample()
    assert d.contains(s) and getattr(d, "shape", None) is None


def test_shard_range_partitions_exactly():
    for total, world in [(16384, 8), (1000, 3), (5, 8), (1024, 1)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_package_imports_without_gpu_and_fails_loudly_for_compute():
    import torch
    import r_mabm_b200
    from r_mabm_b200.engine import BatchedCats
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            BatchedCats(1)
    with pytest.raises(ValueError):
        BatchedCats(1, reward_type="nope")
    assert r_mabm_b200.__version__
