"""Throughput of one build of the library (CATS_B200_LIB) on the configurations that matter.  Development aid."""
import os, sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats
from r_mabm_b200 import _native as nat
lib = nat.load()
tag = os.environ.get("CATS_B200_LIB", "tree").split("/")[-1]

def probe(name, B, periods, warps=0, pipe=-1, burn=300, reps=3, **kw):
    lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
    env = BatchedCats(B, seed=1, **kw)
    env.run(burn)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.run(periods); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"{tag:28s} {name:10s} B={B:5d} warps={warps} pipe={pipe:2d} {best / periods * 1e3:8.1f} us/period {B * periods / (best * 1e-3):.3e}/s un={float(env.scalar('Un').mean()):.6f}", flush=True)
    lib.cats_debug_set_warps(0); lib.cats_debug_set_pipe(-1)

def cold(name, B, burn=300, reps=30, **kw):
    """one period per launch with L2 flushed in between (what bench.py times for configs 2 and 5)"""
    env = BatchedCats(B, seed=1, **kw)
    env.run(burn)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        junk.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.run(1); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort(); med = ts[len(ts) // 2]
    print(f"{tag:28s} {name:10s} B={B:5d} cold 1-period launches: median {med * 1e3:8.1f} us  {B / (med * 1e-3):.3e}/s  info={env.launch_info()}", flush=True)

what = sys.argv[1:] or ["solo", "c2", "s8", "c3", "c5"]
if "solo" in what: probe("solo", 148, 20, warps=1); probe("solo", 148, 20, warps=4, pipe=1)
if "c2" in what: probe("config2", 1024, 50)
if "s8" in what: probe("strong8", 2048, 20)
if "c3" in what: probe("config3", 16384, 5); probe("config3-1p", 16384, 1, reps=10)
if "c2cold" in what: cold("c2cold", 1024)
if "c5cold" in what: cold("c5cold", 256, burn=100, W=10000, F=1000, N=200)
if "c5" in what: probe("config5", 256, 10, burn=100, W=10000, F=1000, N=200)
