"""Per-phase clock64 cycles of economy 0 for the BASELINE configurations (needs a -DCATS_PROFILE_PHASES build:
CATS_B200_LIB=build/libcats_prof.so).  Development aid."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat

NAMES = ["firm", "fire", "job", "produce", "capital", "income+budget", "goods", "accounting", "reward", "tracking",
         "bankrupt", "closing"]


def breakdown(tag, B, periods=20, burn=300, **kw):
    env = BatchedCats(B, seed=1, **kw)
    env.run(burn)
    cyc = torch.zeros(16, dtype=torch.int64, device="cuda")
    nat.load().cats_debug_phase_cycles(cyc.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.run(periods); e.record()
    torch.cuda.synchronize()
    nat.load().cats_debug_phase_cycles(None)
    c = cyc.cpu().numpy()[:12] / periods
    ms = s.elapsed_time(e)
    print(f"{tag}: B={B} {B * periods / ms * 1e3:.3e} econ-periods/s, {ms / periods * 1e3:.1f} us/period; cycles/period of economy 0:",
          {n: int(v) for n, v in zip(NAMES, c)}, "total", int(c.sum()), env.launch_info(), flush=True)


if __name__ == "__main__":
    breakdown("config2", 1024)
    breakdown("config3-strong-8gpu", 2048)
    breakdown("config3", 16384)
    breakdown("solo", 148)
    breakdown("config5", 256, periods=10, burn=100, W=10000, F=1000, N=200)
