"""Times the thread-per-economy goods-market probe (scripts/proto_tpe.cu) on burnt-in economies."""
import ctypes as C, sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

lib = C.CDLL("build/libproto_tpe.so")
lib.proto_goods_launch.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
env = BatchedCats(B, seed=1)
env.run(300)
torch.cuda.synchronize()
# reference point: the whole period with the product kernel
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); env.run(1); e.record(); torch.cuda.synchronize()
print(f"product kernel, full period: {s.elapsed_time(e):.3f} ms")
q_ref = env.field("c_Q").sum().item()
for ilv in (1, 2, 4):
    blob = env.blobs.clone()
    best = 1e9
    for rep in range(3):
        work = blob.clone()
        s.record()
        occ = lib.proto_goods_launch(work.data_ptr(), B, env.W, env.F, env.N, env._params_t.data_ptr(), 1234, ilv, None)
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"TPE goods market ilv={ilv}: {best:.3f} ms  ({B / best / 1e3:.2f}e6 markets/s)  warps/SM={occ}", flush=True)
