import sys, subprocess
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    import torch
    from r_mabm_b200.engine import BatchedCats
    from r_mabm_b200 import _native as nat
    B, warps, pipe, burn, periods, per_launch = (int(x) for x in sys.argv[1:7])
    lib = nat.load()
    env = BatchedCats(B, seed=1)
    lib.cats_debug_set_warps(1)
    env.run(burn) if burn else None
    torch.cuda.synchronize()
    lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
    for i in range(periods):
        env.run(per_launch)
        torch.cuda.synchronize()
    print("ok", float(env.scalar("Un").mean()))
else:
    for args in ((1024, 4, 1, 300, 10, 1), (1024, 4, 1, 0, 5, 10), (1024, 4, 1, 300, 3, 10), (148, 4, 1, 300, 3, 10), (1024, 2, 1, 300, 3, 10),
                 (1024, 8, 1, 300, 3, 10), (1024, 4, 0, 300, 3, 10), (8, 4, 1, 300, 3, 10)):
        r = subprocess.run([sys.executable, __file__] + [str(a) for a in args], capture_output=True, text=True)
        print(args, "rc", r.returncode, (r.stdout.strip().splitlines() or [""])[-1][:100], (r.stderr.strip().splitlines() or [""])[-1][:160], flush=True)
