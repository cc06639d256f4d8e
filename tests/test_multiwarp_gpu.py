"""The multi-warp period kernel (cats_period_mw.cuh: one thread block per economy, 2 / 4 / 8 warps, the wide
goods market with commit rule R2) on the GPU: bit-identical to the oracle and to the one-warp kernel.
`cats_debug_set_warps` forces the shape; left alone the library picks it from the batch size."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import bits, compare_state, make_oracles

pytestmark = pytest.mark.gpu


@pytest.fixture
def lib():
    from r_mabm_b200 import _native as nat
    L = nat.load()
    yield L
    L.cats_debug_set_warps(0)
    L.cats_debug_set_pipe(-1)


def _env(B, **kw):
    from r_mabm_b200.engine import BatchedCats
    return BatchedCats(B, **kw)


@pytest.mark.parametrize("warps,pipe", [(2, 0), (4, 0), (8, 0), (2, 1), (4, 1), (8, 1)])
def test_default_size_120_periods(lib, warps, pipe):
    """pipe 0: the block-wide goods market; 1: the producer / consumer pipeline."""
    lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
    B = 8
    env = _env(B, seed=2024)
    orcs = make_oracles(B, 1000, 100, 20, seed=2024)
    for t in range(0, 120, 20):
        stats = env.run(20, want_stats=True)
        ref = orc.run_batch(orcs, 20, n_threads=8)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"{warps} warps t={t + 20}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))


@pytest.mark.parametrize("W,F,N,zs,warps", [(37, 3, 1, (5, 2, 5), 2), (1, 1, 1, (5, 2, 5), 4), (64, 33, 31, (8, 8, 8), 8),
                                             (130, 7, 2, (2, 1, 3), 4), (500, 50, 10, (4, 3, 6), 8)])
@pytest.mark.parametrize("pipe", [0, 1])
def test_ragged_sizes(lib, W, F, N, zs, warps, pipe):
    lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
    p = dict(orc.DEFAULT_PARAMS, z_c=zs[0], z_k=zs[1], z_e=zs[2])
    env = _env(3, W=W, F=F, N=N, params=dict(p), seed=41)
    orcs = make_oracles(3, W, F, N, params=p, seed=41)
    for t in range(3):
        env.run(15)
        orc.run_batch(orcs, 15, n_threads=3, want_stats=False)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"{(W, F, N)} t={15 * (t + 1)}")
        assert bad == [], "\n".join(bad)


@pytest.mark.parametrize("pipe", [0, 1])
def test_scaled_economy_10x_eight_warps(lib, pipe):
    lib.cats_debug_set_warps(8); lib.cats_debug_set_pipe(pipe)
    B, W, F, N = 3, 10000, 1000, 200
    env = _env(B, W=W, F=F, N=N, seed=51)
    orcs = make_oracles(B, W, F, N, seed=51)
    for t in range(0, 60, 20):
        stats = env.run(20, want_stats=True)
        ref = orc.run_batch(orcs, 20, n_threads=B)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"10x t={t + 20}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))


def test_every_shape_gives_the_same_bits_with_actions_and_heterogeneous_parameters(lib):
    B, A = 37, 2
    plist = [dict(orc.DEFAULT_PARAMS) for _ in range(B)]
    for i, p in enumerate(plist):
        p.update(chi=0.04 + 0.001 * i, Iprob=0.2 + 0.005 * i)
    ref = None
    acts = torch.full((B, A, 2), 1.03, dtype=torch.float64, device="cuda")
    for warps, pipe in ((1, -1), (2, 0), (4, 0), (8, 0), (2, 1), (4, 1), (8, 1), (0, -1)):
        lib.cats_debug_set_warps(warps); lib.cats_debug_set_pipe(pipe)
        env = _env(B, W=400, F=40, N=8, params=[dict(p) for p in plist], n_agents=A, seed=77)
        env.run(25)
        obs, rew, term, st = env.step(acts, want_stats=True)
        stats = env.run(6, want_stats=True)
        torch.cuda.synchronize()
        got = (env.blobs.clone(), obs.clone(), rew.clone(), st.clone(), stats.clone())
        if ref is None:
            ref = got
        else:
            for a, b in zip(got, ref):
                assert torch.equal(a, b), f"{warps} warps, pipe {pipe}"


def test_config2_width_auto_shape_matches_one_warp_kernel(lib):
    """1,024 default-size economies: the library picks the multi-warp kernel by itself; same bits as the
    one-warp kernel, sampled economies equal the oracle."""
    B = 1024
    lib.cats_debug_set_warps(0)
    a = _env(B, seed=5)
    sa = a.run(60, want_stats=True)
    lib.cats_debug_set_warps(1)
    b = _env(B, seed=5)
    sb = b.run(60, want_stats=True)
    torch.cuda.synchronize()
    assert torch.equal(a.blobs, b.blobs) and torch.equal(sa, sb)
    ids = [0, 511, 1023]
    orcs = [orc.OracleEconomy(seed=5, economy=i) for i in ids]
    orc.run_batch(orcs, 60, n_threads=3, want_stats=False)
    for i, o in zip(ids, orcs):
        assert np.array_equal(bits(a.field("PA")[i].cpu().numpy()), bits(o.f64("PA")))
        assert np.array_equal(a.field("Oc")[i].cpu().numpy(), o.i32("Oc"))


def test_launch_shape_keeps_every_block_resident(lib):
    """The library gives an economy several warps only while every block of the launch is resident at once: the
    8-warp kernel holds 3 blocks per SM (registers), the 4-warp kernel 7; beyond that the one-warp kernel."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    lib.cats_debug_set_warps(0); lib.cats_debug_set_pipe(-1)
    for B, want in ((1, 8), (3 * sms, 8), (3 * sms + 1, 4), (7 * sms, 4), (7 * sms + 1, 1), (16384, 1)):
        env = _env(B)
        shape = env.launch_shape()
        assert shape["warps_per_economy"] == want, (B, shape)
        if want > 1:
            assert shape["goods_market"] == "pipelined"
        del env
