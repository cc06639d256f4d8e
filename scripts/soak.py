"""Soak run: B default-size economies for the reference's default episode length (T = 5000 periods after the
300-period burn-in), statistics of every period checked for finiteness and range.  Development aid."""
import sys
sys.path.insert(0, ".")
import torch
from r_mabm_b200.engine import BatchedCats

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedCats(B, seed=123)
env.run(300)
bad = 0
un_sum = 0.0
for chunk in range(20):
    st = env.run(250, want_stats=True)
    torch.cuda.synchronize()
    bad += int((~torch.isfinite(st)).sum())
    un = st[:, :, 2]
    assert ((un >= 0) & (un <= 1)).all(), "unemployment out of range"
    un_sum += float(un.mean())
    if chunk % 5 == 4:
        print(f"t={300 + 250 * (chunk + 1)}: mean Un {float(un[:, -1].mean()):.4f}  mean price {float(st[:, -1, 4].mean()):.4g}  "
              f"mean Y_real {float(st[:, -1, 0].mean()):.4g}  bankruptcy rate {float(st[:, -1, 3].mean()):.4f}  "
              f"status bits {int(env._status.max())}", flush=True)
print(f"non-finite statistics: {bad}; mean unemployment over the run {un_sum / 20:.4f}; "
      f"state finite: {bool(torch.isfinite(env.field('scal')).all())}", flush=True)
