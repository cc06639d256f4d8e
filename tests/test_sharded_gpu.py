"""`ShardedCats` on the device: slabs of a batch give the same economies as the whole batch (the Philox counter
carries the GLOBAL economy id), the gathered statistics reach `Logger.log_batched_stats` in the reference's
per-key `.npy` layout, and -- on a box with two GPUs -- two NCCL ranks reproduce the single-GPU bits."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_slabs_equal_the_whole_batch_and_the_log_files_are_written(tmp_path):
    from r_mabm_b200.engine import BatchedCats
    from r_mabm_b200.logger import Logger
    from r_mabm_b200.sharded import ShardedCats
    from r_mabm_b200._native import STAT_NAMES
    total, T = 11, 12
    whole = BatchedCats(total, W=300, F=30, N=8, seed=5)
    ref = whole.run(T, want_stats=True)
    parts = []
    for rank in range(3):     # uneven slabs: 4 + 4 + 3
        sh = ShardedCats(total, rank=rank, world=3, device="cuda", W=300, F=30, N=8, seed=5)
        st = sh.run(T)
        torch.cuda.synchronize()
        assert torch.equal(sh.env.blobs, whole.blobs[sh.begin:sh.end])
        parts.append(st)
    assert torch.equal(torch.cat(parts, 0), ref)
    # world size 1: run_logged == run + Logger
    solo = ShardedCats(total, rank=0, world=1, device="cuda", W=300, F=30, N=8, seed=5)
    log = Logger("mc", tmp_path, use_timestamp=False)
    full = solo.run_logged(T, logger=log, episode=0, chunk=5)
    assert torch.equal(full, ref)
    un = np.load(tmp_path / "mc" / "ep0_Un.npy")
    assert un.shape == (T, total)
    assert np.array_equal(un, ref[:, :, STAT_NAMES.index("Un")].T.cpu().numpy())
    assert sorted(p.name for p in (tmp_path / "mc").glob("ep0_*.npy")) == sorted(f"ep0_{k}.npy" for k in STAT_NAMES)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rank_main(rank, world, port, total, T, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank); os.environ["WORLD_SIZE"] = str(world); os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from r_mabm_b200.logger import Logger
        from r_mabm_b200.sharded import ShardedCats
        sh = ShardedCats(total, seed=5)                       # rank / world / device from the process group and env
        log = Logger("mc2", out_dir, use_timestamp=False) if rank == 0 else None
        full = sh.run_logged(T, logger=log, episode=0, chunk=7)
        if rank == 0:
            np.save(os.path.join(out_dir, "full.npy"), full.cpu().numpy())
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_over_nccl_equal_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    from r_mabm_b200.engine import BatchedCats
    from r_mabm_b200._native import STAT_NAMES
    total, T, world = 37, 20, 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, total, T, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    whole = BatchedCats(total, seed=5)
    ref = whole.run(T, want_stats=True).cpu().numpy()
    got = np.load(tmp_path / "full.npy")
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    y = np.load(tmp_path / "mc2" / "ep0_Y_real.npy")
    assert np.array_equal(y, ref[:, :, STAT_NAMES.index("Y_real")].T)
