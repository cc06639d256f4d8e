"""`ShardedCats`: one batch of economies over the GPUs of a box, one process per GPU.

Economies never interact (the reference has one economy per process,
/root/reference/rmabm/environments/cats.py:224), so a batch of `total` economies is cut into
contiguous slabs of global economy ids, one per rank; there is NO collective on the step path.
The Philox counter carries the GLOBAL economy id (`economy_base`), so economy g gives the same
bits at any world size -- `tests/test_distributed_cpu.py` (gloo) and the 2-GPU test in
`tests/test_sharded_gpu.py` hold that.  The only communication is the gather of the logged
statistics [B_local, T, 16] at the end of a run (NCCL over NVLink; `gather_stats`), which rank 0
hands to `Logger.log_batched_stats` -- the reference's per-key `.npy` layout
(/root/reference/rmabm/logger.py:72-103, simulation.py:141-146) with an economy axis appended.

Launch with torchrun (RANK / LOCAL_RANK / WORLD_SIZE in the environment) or pass rank / world.
"""
from __future__ import annotations

import os

import torch

from .rollout import gather_stats, shard_range


class ShardedCats:
    """This rank's slab of `total` economies.  Keyword arguments are `BatchedCats`'."""

    def __init__(self, total: int, rank: int | None = None, world: int | None = None, device=None, group=None,
                 **engine_kw):
        import torch.distributed as dist
        if world is None:
            world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else int(os.environ.get("WORLD_SIZE", "1"))
        if rank is None:
            rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else int(os.environ.get("RANK", "0"))
        if total < world:
            raise ValueError(f"{total} economies cannot be sharded over {world} ranks")
        self.total, self.rank, self.world, self.group = int(total), int(rank), int(world), group
        self.begin, self.end = shard_range(self.total, self.rank, self.world)
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", str(self.rank))))
        from .engine import BatchedCats
        self.env = BatchedCats(self.end - self.begin, economy_base=self.begin, device=device, **engine_kw)

    @property
    def local_economies(self) -> int:
        return self.end - self.begin

    def reset(self, seed: int | None = None, t_burnin: int = 300) -> None:
        """`Cats.reset` for every economy of the slab (cats.py:283-315): a new model and `t_burnin` periods."""
        self.env.reset_state(seed)
        if t_burnin > 0:
            self.env.run(t_burnin)

    def run(self, n_periods: int, want_stats: bool = True):
        """Advance the slab `n_periods` heuristic periods in one launch; no communication."""
        return self.env.run(n_periods, want_stats=want_stats)

    def gather(self, local_stats: torch.Tensor) -> torch.Tensor:
        """All ranks' statistics [total, T, 16], in global economy order (the one collective of a run)."""
        return gather_stats(local_stats, self.group)

    def run_logged(self, n_periods: int, logger=None, episode: int | None = None, chunk: int = 250) -> torch.Tensor | None:
        """Run `n_periods` (in launches of `chunk` periods), gather the statistics once and let rank 0 write
        them with `logger.log_batched_stats`.  Returns the gathered [total, T, 16] tensor on rank 0."""
        parts = []
        done = 0
        while done < n_periods:
            n = min(chunk, n_periods - done)
            parts.append(self.env.run(n, want_stats=True))
            done += n
        local = torch.cat(parts, 1) if len(parts) > 1 else parts[0]
        full = self.gather(local)
        if self.rank != 0:
            return None
        if logger is not None:
            logger.log_batched_stats(full.cpu().numpy(), episode=episode)
        return full
