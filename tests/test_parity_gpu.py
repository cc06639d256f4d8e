"""GPU parity: the CUDA period kernel vs the CPU oracle, bit for bit, through the C-ABI.

The oracle is the in-repo CPU restatement of the model spec (docs/MODEL.md) -- "parity unpinned"
with respect to ABCredit.jl, see oracle/cats_oracle.c.  Bar: integer state identical, float64
state identical to the last bit (stronger than the 1e-9 relative tolerance north_star allows).
"""
import numpy as np
import pytest
import torch

from parity_util import compare_state, compare_workset, make_oracles

pytestmark = pytest.mark.gpu


def _env(B, **kw):
    from r_mabm_b200.engine import BatchedCats
    return BatchedCats(B, **kw)


def test_init_matches_oracle():
    env = _env(4, seed=3)
    orcs = make_oracles(4, 1000, 100, 20, seed=3)
    torch.cuda.synchronize()
    assert compare_state(env, orcs, "init") == []


@pytest.mark.parametrize("phase", [3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 24, 25, 26, 27, 28])
def test_phase_by_phase_first_period(phase):
    """Stop the first period after `phase` on both sides; persisted state and transients agree."""
    B = 3
    env = _env(B, seed=11)
    orcs = make_oracles(B, 1000, 100, 20, seed=11)
    # two warm periods so that every market is live, then the probed period
    env.run(2)
    for o in orcs:
        o.step(); o.step()
    dbg = torch.zeros((B, env.workset_bytes()), dtype=torch.uint8, device="cuda")
    env.step(burnin=True, stop_phase=phase, debug=dbg)
    for o in orcs:
        o.set_stop_phase(phase); o.step()
    torch.cuda.synchronize()
    bad = compare_state(env, orcs, f"phase {phase}")
    if phase < 21:  # accounting reuses transient rows as scratch afterwards
        bad += compare_workset(env, dbg, orcs, f"phase {phase}")
    assert bad == [], "\n".join(bad)


def test_100_periods_bit_exact_default_size():
    B, T = 8, 100
    env = _env(B, seed=2024)
    orcs = make_oracles(B, 1000, 100, 20, seed=2024)
    for t in range(0, T, 10):
        stats = env.run(10, want_stats=True)
        from oracle.oracle import run_batch
        ref = run_batch(orcs, 10, n_threads=8)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"t={t + 10}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))


def test_single_step_launches_equal_multi_period_launch():
    a, b = _env(4, seed=5), _env(4, seed=5)
    a.run(20)
    for _ in range(20):
        b.step(burnin=True)
    torch.cuda.synchronize()
    assert torch.equal(a.blobs, b.blobs)


def test_batch_size_and_sharding_invariance():
    """Economy g gives the same bits whether it is blob g of one launch or blob 0 of a shard."""
    full = _env(6, seed=9)
    full.run(15)
    part = _env(2, seed=9, economy_base=3)
    part.run(15)
    torch.cuda.synchronize()
    assert torch.equal(full.blobs[3:5], part.blobs)


def test_block_shape_never_changes_results():
    """Economies per block is a scheduling choice: 1, 2, 4, 8, 12 and 14 per block, with
    a ragged last block and with the per-period barriers of a multi-period launch, give the same bits."""
    from r_mabm_b200 import _native as nat
    from r_mabm_b200.engine import BatchedCats
    lib = nat.load()
    ref = None
    try:
        for g in (1, 2, 4, 8, 12, 14, 0):
            lib.cats_debug_set_group(g)
            env = BatchedCats(29, W=200, F=20, N=5, n_agents=1, seed=11)
            env.run(7)
            acts = torch.full((29, 1, 2), 1.02, dtype=torch.float64, device="cuda")
            env.step(acts, torch.ones((29, 1), dtype=torch.uint8, device="cuda"))
            stats = env.run(5, want_stats=True)
            torch.cuda.synchronize()
            got = (env.blobs.clone(), stats.clone())
            if ref is None:
                ref = got
            else:
                assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), f"group {g}"
    finally:
        lib.cats_debug_set_group(0)


def test_static_layout_kernel_equals_the_generic_kernel_at_the_default_size():
    """Default-size economies run a kernel whose layout is a compile-time constant.
    The generic instantiation, every block shape, RL actions, statistics and a masked launch (always
    generic) give the same bits."""
    from r_mabm_b200 import _native as nat
    from r_mabm_b200.engine import BatchedCats
    lib = nat.load()
    ref = None
    try:
        for static, g in ((1, 0), (0, 0), (1, 8), (0, 8), (1, 1), (1, 12)):
            lib.cats_debug_set_static(static); lib.cats_debug_set_group(g)
            env = BatchedCats(31, n_agents=2, seed=19)          # W=1000, F=100, N=20
            env.run(9)
            acts = torch.full((31, 2, 2), 0.97, dtype=torch.float64, device="cuda")
            obs, rew, term, st1 = env.step(acts, torch.ones((31, 2), dtype=torch.uint8, device="cuda"), want_stats=True)
            obs, rew = obs.clone(), rew.clone()
            stats = env.run(4, want_stats=True)
            env.run(2, active=torch.ones(31, dtype=torch.uint8, device="cuda"))
            torch.cuda.synchronize()
            got = (env.blobs.clone(), stats.clone(), obs, rew, st1.clone())
            if ref is None:
                ref = got
            else:
                for a, b in zip(got, ref):
                    assert torch.equal(a, b), f"static={static} group={g}"
    finally:
        lib.cats_debug_set_static(1); lib.cats_debug_set_group(0)


def test_default_size_with_other_z_takes_the_generic_kernel_and_a_false_hint_is_flagged():
    """The default-configuration kernel has z_c, z_k, z_e = 5, 2, 5 compiled in.  Other values at the
    default size run the generic production kernel (checked against the oracle); a call that declares
    the defaults while an economy's parameters say otherwise gets CATS_STATUS_Z_MISMATCH."""
    import ctypes as C
    from r_mabm_b200 import _native as nat
    from r_mabm_b200.engine import BatchedCats
    from oracle import oracle as orc
    from parity_util import compare_state
    p = {"z_c": 4, "z_k": 3, "z_e": 2}
    env = BatchedCats(3, params=dict(p), seed=5)
    env.run(6)
    torch.cuda.synchronize()
    refs = [orc.OracleEconomy(1000, 100, 20, params=dict(p), seed=5, economy=b) for b in range(3)]
    for o in refs:
        for _ in range(6):
            o.step()
    assert compare_state(env, refs, "z = 4, 3, 2") == []
    assert int(env._status.max()) == 0
    a = env._args(1, env.flags | nat.FLAG_BURNIN, want_io=False)
    a.z_all[0], a.z_all[1], a.z_all[2] = 5, 2, 5            # a lie
    nat.check(env._lib.cats_step(C.byref(a)))
    torch.cuda.synchronize()
    assert (env._status.cpu().numpy() & nat.STATUS_Z_MISMATCH).all()
