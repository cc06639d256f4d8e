"""BASELINE.json config 4 (MARL rollout): 4,096 CatsLog economies, 2 learned firms each, env step + tabular-Q
policy + TD update on the device.  Prints rollout steps/s and where the time goes (development aid)."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200.environments import CatsLog
from r_mabm_b200.rollout import BatchedQPolicy, BatchedRollout, FusedQPolicy

B, A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 2
env = BatchedCats(B, n_agents=A, log_mode=True, seed=1)
pol = BatchedQPolicy(B, A, dict(CatsLog._default_gym_spaces_bounds), epsilon=0.3, device="cuda", seed=2)
ro = BatchedRollout(env, pol, t_burnin=300, T=1000)
ro.reset(seed=1)


def timed(fn, n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for _ in range(20):
    ro.step(train=True)
ms_full = timed(lambda: ro.step(train=True), 200)
acts = pol.get_action(ro.obs)
ms_env = timed(lambda: env.step(acts), 200)
ms_act = timed(lambda: pol.get_action(ro.obs), 200)
obs, rew, term, _ = env.step(acts)
ms_train = timed(lambda: pol.train(rew, obs, term), 200)
print(f"B={B}: rollout step {ms_full:.3f} ms = {B / ms_full * 1e3:.3e} economy-periods/s; "
      f"env.step {ms_env:.3f} ms, policy.get_action {ms_act:.3f} ms, policy.train {ms_train:.3f} ms", flush=True)
fpol = FusedQPolicy(B, A, dict(CatsLog._default_gym_spaces_bounds), epsilon=0.3, seed=2)
fro = BatchedRollout(env, fpol, t_burnin=300, T=1000)
fro.reset(seed=1)
for _ in range(20):
    fro.step(train=True)
ms_fused = timed(lambda: fro.step(train=True), 300)
ms_q = timed(lambda: fpol.train_and_act(rew, obs, term), 300)
print(f"B={B}: fused policy kernel: rollout step {ms_fused:.3f} ms = {B / ms_fused * 1e3:.3e} economy-periods/s "
      f"(policy kernel alone {ms_q * 1e3:.1f} us)", flush=True)
