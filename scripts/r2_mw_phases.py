"""Per-phase cycles of economy 0 in the multi-warp kernel (needs CATS_B200_LIB=build/libcats_prof.so)."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat

NAMES = ["firm", "fire", "job", "produce", "capital", "gov", "goods", "accounting", "reward", "tracking", "bankrupt",
         "closing+sync", "g.prep", "g.walk+pub", "g.replay+cascade", "g.commit+reset"]
lib = nat.load()


def breakdown(tag, B, warps, periods=20, burn=300, **kw):
    lib.cats_debug_set_warps(warps)
    env = BatchedCats(B, seed=1, **kw)
    env.run(burn)
    cyc = torch.zeros(16, dtype=torch.int64, device="cuda")
    lib.cats_debug_phase_cycles(cyc.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.run(periods); e.record()
    torch.cuda.synchronize()
    lib.cats_debug_phase_cycles(None)
    c = cyc.cpu().numpy() / periods
    ms = s.elapsed_time(e)
    print(f"{tag} warps={warps}: B={B} {ms / periods * 1e3:.1f} us/period;", {n: int(v) for n, v in zip(NAMES, c)}, flush=True)
    lib.cats_debug_set_warps(0)


if __name__ == "__main__":
    for w in (1, 2, 4, 8):
        breakdown("solo", 148, w)
    for w in (1, 4):
        breakdown("config2", 1024, w)
    for w in (1, 8):
        breakdown("config5", 256, w, periods=10, burn=100, W=10000, F=1000, N=200)
