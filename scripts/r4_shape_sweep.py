"""Throughput of every kernel shape (warps per economy x goods-market variant) over launch widths.  Development aid."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat
lib = nat.load()

def probe(B, periods, warps, pipe, burn=300, reps=3):
    lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
    try:
        env = BatchedCats(B, seed=1)
        env.run(burn)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.run(periods); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        print(f"B={B:5d} warps={warps} pipe={pipe:2d} {best / periods * 1e3:8.1f} us/period {B * periods / (best * 1e-3):.3e}/s un={float(env.scalar('Un').mean()):.6f}", flush=True)
    except Exception as ex:
        print(f"B={B} warps={warps} pipe={pipe}: {type(ex).__name__}: {ex}", flush=True)
    lib.cats_debug_set_warps(0); lib.cats_debug_set_pipe(-1)

Bs = [int(x) for x in sys.argv[1:]] or [1024, 2048, 3072, 4096]
for B in Bs:
    for warps, pipe in ((1, -1), (2, 1), (4, 1), (8, 1)):
        probe(B, 20, warps, pipe)
