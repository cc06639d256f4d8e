"""Quick device-side throughput probe (development aid, not the bench contract).
usage: quick_perf.py [B periods]..."""
import sys, time
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

def probe(B, periods, reps=3, **kw):
    env = BatchedCats(B, seed=1, **kw)
    env.run(300)  # the reference burn-in: past the start-up transient
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.run(periods); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    rate = B * periods / (best * 1e-3)
    print(f"B={B} periods/launch={periods} ms={best:.2f} econ-periods/s={rate:.3e} info={env.launch_info()}", flush=True)
    return rate

def phase_breakdown(B=148, periods=20):
    from r_mabm_b200 import _native as nat
    env = BatchedCats(B, seed=1)
    env.run(50)
    cyc = torch.zeros(16, dtype=torch.int64, device="cuda")
    nat.load().cats_debug_phase_cycles(cyc.data_ptr())
    env.run(periods)
    torch.cuda.synchronize()
    nat.load().cats_debug_phase_cycles(None)
    names = ["firm", "fire", "job", "produce", "capital", "income+budget", "goods", "accounting", "reward", "tracking", "bankrupt", "closing"]
    c = cyc.cpu().numpy()[:12] / periods
    print("cycles/period:", {n: int(v) for n, v in zip(names, c)}, "total", int(c.sum()), flush=True)


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    if a:
        for i in range(0, len(a), 2):
            probe(a[i], a[i + 1])
    else:
        for B in (148, 1024, 4096):
            probe(B, 20)
        probe(1024, 1)
        probe(16384, 1)
        phase_breakdown()
