"""Small, short launch for ncu captures: B economies after the 300-period burn-in, then `periods` periods in one launch."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
periods = int(sys.argv[2]) if len(sys.argv) > 2 else 1
env = BatchedCats(B, seed=1)
env.run(300)         # kernel launch 1 (after init): burn-in
env.run(periods)     # kernel launch 2: the one to profile
torch.cuda.synchronize()
print("ok", float(env.scalar("Un").mean()))
