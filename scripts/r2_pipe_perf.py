"""Throughput per kernel shape (warps per economy x goods-market variant).  Development aid."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat

lib = nat.load()


def probe(tag, B, periods, shapes, burn=300, reps=3, **kw):
    for warps, pipe in shapes:
        lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
        try:
            env = BatchedCats(B, seed=1, **kw)
            env.run(burn)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(reps):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); env.run(periods); e.record(); torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e))
            print(f"{tag}: B={B} warps={warps} pipe={pipe} periods/launch={periods} {best / periods * 1e3:.1f} us/period "
                  f"{B * periods / (best * 1e-3):.3e} econ-periods/s", flush=True)
        except Exception as ex:
            print(f"{tag}: B={B} warps={warps} pipe={pipe}: {type(ex).__name__}: {ex}", flush=True)
    lib.cats_debug_set_warps(0); lib.cats_debug_set_pipe(-1)


if __name__ == "__main__":
    S = [(1, -1), (2, 1), (4, 1), (8, 1), (4, 0), (8, 0), (0, -1)]
    probe("solo", 148, 20, S)
    probe("config2", 1024, 50, [(1, -1), (2, 1), (4, 1), (4, 0), (0, -1)])
    probe("config2-1period", 1024, 1, [(1, -1), (4, 1), (0, -1)], reps=20)
    probe("strong8", 2048, 20, [(1, -1), (2, 1), (4, 1), (0, -1)])
    probe("strong4", 4096, 20, [(1, -1), (2, 1), (0, -1)])
    probe("config3", 16384, 5, [(1, -1), (2, 1), (0, -1)])
    probe("config5", 256, 10, [(1, -1), (8, 1), (8, 0), (0, -1)], burn=100, W=10000, F=1000, N=200)
