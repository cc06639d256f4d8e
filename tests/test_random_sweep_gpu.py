"""GPU: seeded random sweep over economy sizes, candidate-list lengths, parameter values, reward / price /
log modes, action masks and heterogeneous parameters -- every case stepped on the GPU and by the oracle and
compared to the bit (state, observations, rewards).  Complements the hand-picked cases of
test_parity_gpu*.py: sizes that are not multiples of 32, F < z, N = 1, a single worker per firm, ..."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import bits, compare_state, make_oracles

pytestmark = pytest.mark.gpu


def random_case(rng):
    F = int(rng.integers(1, 70))
    N = int(rng.integers(1, 14))
    W = int(rng.integers(max(F + N, 8), 700))
    base = {"z_c": int(rng.integers(1, 9)), "z_k": int(rng.integers(1, 9)), "z_e": int(rng.integers(1, 9)),
            "xi": float(rng.uniform(0.8, 0.99)), "chi": float(rng.uniform(0.01, 0.2)),
            "q_adj": float(rng.uniform(0.5, 1.0)), "p_adj": float(rng.uniform(0.02, 0.3)),
            "Iprob": float(rng.uniform(0.05, 0.6)), "theta": float(rng.uniform(0.02, 0.2)),
            "tax_rate": float(rng.choice([0.0, 0.1])), "subsidy": float(rng.choice([0.0, 0.3])),
            "tax_rate_d": float(rng.choice([0.0, 0.15])), "maastricht": float(rng.choice([0.0, 1.0])),
            "inventory_depreciation": float(rng.uniform(0.0, 0.6)), "wage_update_up": float(rng.uniform(0.0, 0.2))}
    return W, F, N, base


@pytest.mark.parametrize("case", range(64))
def test_random_economies_match_the_oracle(case):
    from r_mabm_b200.engine import BatchedCats
    rng = np.random.default_rng(1000 + case)
    W, F, N, base = random_case(rng)
    B = 3
    hetero = case % 3 == 0
    plist = []
    for b in range(B):
        p = dict(base)
        if hetero:
            p["p_adj"] = float(rng.uniform(0.02, 0.3)); p["z_c"] = int(rng.integers(1, 9))
        plist.append(p)
    A = int(rng.integers(0, min(F, 3) + 1))
    log_mode = bool(rng.integers(0, 2))
    price_change = "avg_price" if rng.integers(0, 2) else "agent_price"
    reward_type = "rms" if rng.integers(0, 2) else "profits"
    seed, base_id = int(rng.integers(0, 2**40)), int(rng.integers(0, 10**6))
    env = BatchedCats(B, W=W, F=F, N=N, params=[dict(p) for p in plist] if hetero else dict(base), n_agents=A,
                      log_mode=log_mode, price_change=price_change, reward_type=reward_type, seed=seed,
                      economy_base=base_id, bankruptcy_reward=-3.0)
    refs = make_oracles(B, W, F, N, params=[dict(p) for p in plist] if hetero else dict(base), seed=seed,
                        economy_base=base_id, bankruptcy_reward=-3.0)
    burn = int(rng.integers(1, 9))
    env.run(burn)
    for o in refs:
        for _ in range(burn):
            o.step()
    fl = ((orc.FLAG_LOG if log_mode else 0) | (orc.FLAG_AVG_PRICE if price_change == "avg_price" else 0)
          | (orc.FLAG_REWARD_RMS if reward_type == "rms" else 0))
    what = f"case {case}: W={W} F={F} N={N} A={A} z={base['z_c']},{base['z_k']},{base['z_e']} log={log_mode} {price_change} {reward_type}"
    for t in range(5):
        if A:
            lo, hi = (-0.3, 0.3) if log_mode else (0.7, 1.3)
            acts = rng.uniform(lo, hi, size=(B, A, 2))
            mask = (rng.random((B, A)) < 0.7).astype(np.uint8)
            obs, rew, term, _ = env.step(torch.from_numpy(acts).cuda(), torch.from_numpy(mask).cuda())
            torch.cuda.synchronize()
            g_obs, g_rew = obs.cpu().numpy(), rew.cpu().numpy()
            for b, o in enumerate(refs):
                r_obs, r_rew = o.step(actions=acts[b], mask=mask[b], flags=fl)
                assert np.array_equal(bits(g_obs[b]), bits(r_obs)), (what, t, b)
                assert np.array_equal(bits(g_rew[b]), bits(r_rew)), (what, t, b)
        else:
            env.run(1)
            for o in refs:
                o.step()
        torch.cuda.synchronize()
        assert compare_state(env, refs, f"{what} t={t}") == []
