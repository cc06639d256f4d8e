"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the path (economy sharding by global id
and the final statistics gather).  There is no collective on the step path itself."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from r_mabm_b200.rollout import gather_stats, shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        lo, hi = shard_range(total, rank, world)
        # every rank advances ITS slab of economies (global ids) -- here with the CPU oracle
        econs = [orc.OracleEconomy(60, 6, 3, seed=4, economy=g) for g in range(lo, hi)]
        stats = orc.run_batch(econs, 5, n_threads=1)
        local = torch.from_numpy(stats[:, -1, :].copy())
        allst = gather_stats(local)
        if rank == 0:
            q.put(allst.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_run_equals_single_process():
    total, world = 5, 2   # uneven slabs: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import oracle as orc
    econs = [orc.OracleEconomy(60, 6, 3, seed=4, economy=g) for g in range(total)]
    ref = orc.run_batch(econs, 5, n_threads=1)[:, -1, :]
    assert got.shape == (total, 16)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))


def test_gather_stats_without_process_group_is_identity():
    x = torch.arange(6.0).view(3, 2)
    assert gather_stats(x) is x
