"""Per-phase clock64 totals and goods-market loop counts of economy 0 (needs a -DCATS_PROFILE_PHASES build)."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3552
periods = 20
env = BatchedCats(B, seed=1)
env.run(300)
cyc = torch.zeros(16, dtype=torch.int64, device="cuda")
nat.load().cats_debug_phase_cycles(cyc.data_ptr())
env.run(periods)
torch.cuda.synchronize()
nat.load().cats_debug_phase_cycles(None)
names = ["firm", "fire", "job", "produce", "capital", "income+budget", "goods", "accounting", "reward", "tracking", "bankrupt", "closing"]
c = cyc.cpu().numpy() / periods
print("cycles/period:", {n: int(v) for n, v in zip(names, c[:12])}, "total", int(c[:12].sum()))
print(f"goods market cycles per period: prepare {c[12]:.0f}, walk {c[13]:.0f}, match+replay {c[14]:.0f}, commit {c[15]:.0f}")
