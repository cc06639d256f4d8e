"""Config 5 (256 economies of 10x size): every kernel shape.  Development aid."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat
lib = nat.load()
for warps, pipe in ((1, -1), (4, 0), (4, 1), (8, 0), (8, 1)):
    lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
    try:
        env = BatchedCats(256, W=10000, F=1000, N=200, seed=1)
        env.run(100)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.run(10); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        print(f"10x B=256 warps={warps} pipe={pipe:2d} {best / 10 * 1e3:8.1f} us/period {256 * 10 / (best * 1e-3):.3e}/s un={float(env.scalar('Un').mean()):.6f}", flush=True)
    except Exception as ex:
        print(f"warps={warps} pipe={pipe}: {type(ex).__name__}: {ex}", flush=True)
lib.cats_debug_set_warps(0); lib.cats_debug_set_pipe(-1)
