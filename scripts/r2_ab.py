"""A/B of two builds on one box: config-2 and config-5 rates (development aid)."""
import subprocess, sys, os
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import torch
    from r_mabm_b200.engine import BatchedCats
    for (tag, B, periods, kw, burn) in (("config2", 1024, 50, {}, 300), ("config5", 256, 10, dict(W=10000, F=1000, N=200), 100)):
        env = BatchedCats(B, seed=1, **kw); env.run(burn); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.run(periods); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        print(f"  {tag}: {best / periods * 1e3:.1f} us/period {B * periods / (best * 1e-3):.3e}", flush=True)
else:
    for rep in range(2):
        for name, lib in (("A(HEAD)", "build/libcats_A.so"), ("B(tree)", "r_mabm_b200/libcats_b200.so")):
            print(name, flush=True)
            env = dict(os.environ, CATS_B200_LIB=lib)
            subprocess.run([sys.executable, __file__, "child"], env=env)
