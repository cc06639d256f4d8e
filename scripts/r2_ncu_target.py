"""Short launches for ncu captures (round 2).
  period  B periods [W F N [warps]]   B economies after the burn-in, then `periods` periods in ONE launch (the one to profile)
  tape    B                           one replay-tape period (the MODE 0 kernel) at the default size
  rollout B steps                     config 4: CatsLog, 2 learned firms, period kernel + fused policy kernel per step
"""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import torch
from r_mabm_b200 import _native as nat
from r_mabm_b200.engine import BatchedCats

kind = sys.argv[1]
if kind == "period":
    B, periods = int(sys.argv[2]), int(sys.argv[3])
    W, F, N = (int(x) for x in sys.argv[4:7]) if len(sys.argv) > 6 else (1000, 100, 20)
    warps = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    nat.load().cats_debug_set_warps(warps)
    env = BatchedCats(B, W=W, F=F, N=N, seed=1)
    env.run(300 if W <= 1000 else 100)   # launch 1 (after init): burn-in
    env.run(periods)                     # launch 2: the one to profile
    torch.cuda.synchronize()
    print("ok", float(env.scalar("Un").mean()), env.launch_info())
elif kind == "tape":
    from tape_util import foreign_tapes
    B = int(sys.argv[2])
    env = BatchedCats(B, seed=1)
    env.run(300)
    recs = foreign_tapes(np.random.default_rng(1), 8, 1000, 100, 20, 5, 2, 5)
    tape = torch.from_numpy(recs).cuda()[torch.arange(B, device="cuda") % 8].contiguous()
    env.run(1, tape=tape.view(-1))       # the launch to profile (MODE 0)
    torch.cuda.synchronize()
    print("ok", float(env.scalar("Un").mean()))
elif kind == "rollout":
    from r_mabm_b200.environments import CatsLog
    from r_mabm_b200.rollout import BatchedRollout, FusedQPolicy
    B, steps = int(sys.argv[2]), int(sys.argv[3])
    env = BatchedCats(B, n_agents=2, log_mode=True, seed=1)
    pol = FusedQPolicy(B, 2, dict(CatsLog._default_gym_spaces_bounds), epsilon=0.3, seed=2)
    ro = BatchedRollout(env, pol, t_burnin=300, T=10 ** 9)
    ro.reset(seed=1)
    for _ in range(steps):
        ro.step(train=True)
    torch.cuda.synchronize()
    print("ok", float(env.scalar("Un").mean()))
