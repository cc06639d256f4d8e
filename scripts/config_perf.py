"""Device-side throughput of the BASELINE.json configurations other than the bench line (development aid)."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

def probe(name, B, periods, reps=3, burn=300, **kw):
    env = BatchedCats(B, seed=1, **kw)
    env.run(burn)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.run(periods); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"{name}: B={B} periods/launch={periods} ms={best:.2f} econ-periods/s={B * periods / (best * 1e-3):.3e} "
          f"launch={env.launch_info()} blob={env.blob_bytes}", flush=True)

probe("config2 1024 x 1000 periods (250/launch)", 1024, 250)
probe("config3 16384 x 1 period", 16384, 1)
probe("config3 16384 x 20 periods/launch", 16384, 20)
probe("config5 10x economy, 256", 256, 10, burn=100, W=10000, F=1000, N=200)
probe("config5 10x economy, 1024", 1024, 10, burn=100, W=10000, F=1000, N=200)
