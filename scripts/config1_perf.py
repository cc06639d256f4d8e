"""BASELINE.json config 1 (examples/heuristic_run.py): ONE default-size economy through the reference-shaped
`Cats` adapter, one Dummy firm, info_level as given.  Steps per second of the Python-facing loop."""
import sys, time
sys.path.insert(0, ".")
from r_mabm_b200.environments import Cats
from r_mabm_b200.agents import Dummy

for lvl in (0, 1, 2):
    env = Cats(T=5000, n_agents=1, info_level=lvl)
    agent = Dummy(env.agents_ids[0], env)
    t0 = time.perf_counter(); obs, info = env.reset(seed=1); t_reset = time.perf_counter() - t0
    n = 1000
    t0 = time.perf_counter()
    for _ in range(n):
        obs, rew, term, trunc, info = env.step([agent.get_action(obs[0])])
    dt = time.perf_counter() - t0
    print(f"info_level={lvl}: reset (init + 300 burn-in periods) {t_reset * 1e3:.1f} ms; step {dt / n * 1e6:.0f} us = {n / dt:.0f} steps/s", flush=True)
