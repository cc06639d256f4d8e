import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
for B in (256, 1024):
    env = BatchedCats(B, seed=1, W=10000, F=1000, N=200)
    env.run(50); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.run(10); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(B, f"{B * 10 / (best * 1e-3):.3e}", env.launch_info(), flush=True)
