"""Throughput of the strict-arithmetic mode (CATS_FLAG_STRICT, MODEL.md 4.6) beside the fast mode.  Development aid."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

for B in (16384, 1024):
    for strict in (False, True):
        env = BatchedCats(B, seed=1, strict=strict)
        env.run(300)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                env.run(1)
            e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) / 20)
        print(f"B={B} strict={strict}: {best * 1e3:.1f} us/period {B / (best * 1e-3):.3e} economy-periods/s", flush=True)
