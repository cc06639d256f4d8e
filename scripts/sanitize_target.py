"""Small, short run for compute-sanitizer: a few economies, every entry point once."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

env = BatchedCats(30, W=120, F=12, N=4, n_agents=2, seed=3)
env.run(12)
acts = torch.full((30, 2, 2), 1.05, dtype=torch.float64, device="cuda")
mask = torch.ones((30, 2), dtype=torch.uint8, device="cuda")
for _ in range(3):
    env.step(acts, mask, want_stats=True)
big = BatchedCats(26, seed=5)          # default size, two blocks of 12 + a ragged one
big.run(3)
torch.cuda.synchronize()
print("ok", float(env.scalar("Un").mean()), float(big.scalar("Un").mean()))
