"""Small, short launch for ncu captures: B economies, a few multi-period launches."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
periods = int(sys.argv[2]) if len(sys.argv) > 2 else 4
env = BatchedCats(B, seed=1)
env.run(60)          # launch 2 (after init): warm-up past the start-up transient
env.run(periods)     # launch 3: the one to profile
torch.cuda.synchronize()
print("ok", float(env.scalar("Un").mean()))
