"""The period kernel's SOURCE, compiled for a CPU emulation of warps and blocks (tests/simt), against the
oracle -- so the kernel's logic is held to the oracle in the CPU suite too, not only on the GPU box, and
under several lane schedules: a result that depends on the order in which the lanes of a warp run between
two collectives is a missing __syncwarp / barrier (the race check compute-sanitizer would do on a GPU).

This is test infrastructure: nothing in r_mabm_b200/ can run on a CPU.  The GPU parity tests
(tests/test_*gpu*.py, through the C-ABI) remain the parity tests proper.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from parity_util import bits, compare_state, compare_workset, make_oracles
from simt.emu import EmuCats
from tape_util import foreign_tapes


@pytest.mark.parametrize("mode,g,schedule", [(2, 1, 0), (2, 2, 1), (1, 2, 5), (0, 1, 1)])
def test_default_size_kernels_match_oracle(mode, g, schedule):
    B, T = 3, 24
    env = EmuCats(B, seed=21, mode=mode, g=g, schedule=schedule)
    orcs = make_oracles(B, 1000, 100, 20, seed=21)
    for t0 in range(0, T, 8):
        stats = env.run(8, want_stats=True)
        ref = orc.run_batch(orcs, 8, n_threads=B)
        bad = compare_state(env, orcs, f"mode {mode} t={t0 + 8}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.view(np.uint64), ref.view(np.uint64))


@pytest.mark.parametrize("W,F,N,zs", [(37, 3, 1, (5, 2, 5)), (1, 1, 1, (5, 2, 5)), (64, 33, 31, (8, 8, 8)),
                                       (130, 7, 2, (2, 1, 3)), (400, 40, 8, (5, 2, 5))])
def test_ragged_sizes_and_lane_schedules_agree(W, F, N, zs):
    p = dict(orc.DEFAULT_PARAMS, z_c=zs[0], z_k=zs[1], z_e=zs[2])
    orcs = make_oracles(3, W, F, N, params=p, seed=41)
    orc.run_batch(orcs, 30, n_threads=3, want_stats=False)
    for schedule, g in ((0, 1), (1, 2), (9, 2)):
        env = EmuCats(3, W=W, F=F, N=N, params=dict(p), seed=41, mode=1, g=g, schedule=schedule)
        env.run(30)
        bad = compare_state(env, orcs, f"{(W, F, N)} schedule {schedule}")
        assert bad == [], "\n".join(bad)


@pytest.mark.parametrize("phase", [10, 12, 13, 16, 19, 20, 22, 25])
def test_phase_by_phase(phase):
    B, W, F, N = 2, 300, 30, 8
    env = EmuCats(B, W=W, F=F, N=N, seed=11, mode=0)
    orcs = make_oracles(B, W, F, N, seed=11)
    env.run(3)
    orc.run_batch(orcs, 3, want_stats=False)
    dbg = np.zeros((B, env.workset_bytes()), dtype=np.uint8)
    env.step(burnin=True, stop_phase=phase, debug=dbg)
    for o in orcs:
        o.set_stop_phase(phase); o.step()
    import torch
    bad = compare_state(env, orcs, f"phase {phase}")
    if phase < 21:
        bad += compare_workset(env, torch.from_numpy(dbg), orcs, f"phase {phase}")
    assert bad == [], "\n".join(bad)


def test_rl_actions_masks_log_mode_and_rewards():
    B, A, W, F, N = 3, 3, 300, 30, 8
    for flags in (0, orc.FLAG_LOG | orc.FLAG_AVG_PRICE | orc.FLAG_REWARD_RMS):
        env = EmuCats(B, W=W, F=F, N=N, n_agents=A, seed=17, mode=1, g=2, flags=flags, bankruptcy_reward=-7.5)
        orcs = make_oracles(B, W, F, N, seed=17, bankruptcy_reward=-7.5)
        env.run(20)
        orc.run_batch(orcs, 20, want_stats=False)
        rng = np.random.default_rng(5)
        for t in range(6):
            lo, hi = (-0.3, 0.3) if flags else (0.7, 1.3)
            acts = rng.uniform(lo, hi, size=(B, A, 2))
            mask = (rng.random((B, A)) < 0.7).astype(np.uint8)
            obs, rew, term, _ = env.step(acts, mask)
            for b, o in enumerate(orcs):
                r_obs, r_rew = o.step(actions=acts[b], mask=mask[b], flags=flags)
                assert np.array_equal(bits(obs[b]), bits(r_obs)) and np.array_equal(bits(rew[b]), bits(r_rew))
            assert compare_state(env, orcs, f"t={t}") == []


def test_foreign_tape_and_strict_mode():
    B, W, F, N = 2, 300, 30, 8
    rng = np.random.default_rng(3)
    for strict in (False, True):
        env = EmuCats(B, W=W, F=F, N=N, seed=5, mode=0, strict=strict, schedule=1)
        orcs = make_oracles(B, W, F, N, seed=5)
        fl = orc.FLAG_STRICT if strict else 0
        for t0 in range(0, 24, 6):
            recs = foreign_tapes(rng, B * 6, W, F, N, 5, 2, 5).reshape(B, 6, -1)
            env.run(6, tape=recs)
            orc.run_batch(orcs, 6, tape=recs, fill_tape=False, flags=fl, want_stats=False)
            bad = compare_state(env, orcs, f"strict={strict} t={t0 + 6}")
            assert bad == [], "\n".join(bad)


@pytest.mark.parametrize("mode", [1, 2])
def test_strict_mode_philox(mode):
    W, F, N = (1000, 100, 20) if mode == 2 else (200, 20, 5)
    env = EmuCats(2, W=W, F=F, N=N, seed=9, mode=mode, strict=True)
    orcs = make_oracles(2, W, F, N, seed=9)
    env.run(16)
    orc.run_batch(orcs, 16, flags=orc.FLAG_STRICT, want_stats=False)
    assert compare_state(env, orcs, "strict") == []


def test_masked_reset_and_episode_stream():
    B, W, F, N = 4, 200, 20, 5
    env = EmuCats(B, W=W, F=F, N=N, seed=31, mode=1, g=2)
    orcs = make_oracles(B, W, F, N, seed=31)
    env.run(12)
    orc.run_batch(orcs, 12, want_stats=False)
    act = np.array([1, 0, 0, 1], dtype=np.uint8)
    env.reset_where(act)
    env.run(9, active=act)
    for b in (0, 3):
        orcs[b].reset()
        for _ in range(9):
            orcs[b].step()
    assert compare_state(env, orcs, "after masked reset") == []
    env.run(5)
    orc.run_batch(orcs, 5, want_stats=False)
    assert compare_state(env, orcs, "after masked reset + 5") == []


@pytest.mark.parametrize("W,F,N,warps,mode,schedule", [(1000, 100, 20, 4, 2, 0), (1000, 100, 20, 8, 1, 3), (1000, 100, 20, 2, 2, 1),
                                                         (300, 30, 8, 4, 1, 1), (130, 7, 2, 8, 1, 6), (37, 3, 1, 2, 1, 0),
                                                         (2000, 200, 40, 8, 1, 1), (900, 300, 12, 4, 1, 4)])
def test_multiwarp_kernel_matches_oracle(W, F, N, warps, mode, schedule):
    """cats_period_mw.cuh: one block per economy, `warps` warps, the wide goods market with commit rule R2."""
    B, T = 2, 60
    env = EmuCats(B, W=W, F=F, N=N, seed=23, mode=mode, warps=warps, schedule=schedule)
    orcs = make_oracles(B, W, F, N, seed=23)
    for t0 in range(0, T, 20):
        stats = env.run(20, want_stats=True)
        ref = orc.run_batch(orcs, 20, n_threads=B)
        bad = compare_state(env, orcs, f"{warps} warps t={t0 + 20}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.view(np.uint64), ref.view(np.uint64))


def test_multiwarp_rl_step():
    B, A, W, F, N = 2, 2, 300, 30, 8
    env = EmuCats(B, W=W, F=F, N=N, n_agents=A, seed=17, mode=1, warps=4, schedule=1)
    orcs = make_oracles(B, W, F, N, seed=17)
    env.run(20)
    orc.run_batch(orcs, 20, want_stats=False)
    rng = np.random.default_rng(5)
    for t in range(5):
        acts = rng.uniform(0.7, 1.3, size=(B, A, 2))
        obs, rew, term, _ = env.step(acts)
        for b, o in enumerate(orcs):
            r_obs, r_rew = o.step(actions=acts[b], flags=0)
            assert np.array_equal(bits(obs[b]), bits(r_obs)) and np.array_equal(bits(rew[b]), bits(r_rew))
        assert compare_state(env, orcs, f"t={t}") == []


@pytest.mark.parametrize("W,F,N,warps,mode,schedule", [(1000, 100, 20, 2, 2, 0), (1000, 100, 20, 4, 2, 3), (1000, 100, 20, 8, 1, 1),
                                                         (300, 30, 8, 4, 1, 5), (130, 7, 2, 8, 1, 6), (37, 3, 1, 2, 1, 1)])
def test_multiwarp_pipelined_goods_market(W, F, N, warps, mode, schedule):
    """mw_goods_market_pc: one warp commits with the one-warp kernel's rounds, the others prepare households into a
    ring of shared-memory slots (mbarrier hand-over, two slots per producer)."""
    B, T = 2, 60
    env = EmuCats(B, W=W, F=F, N=N, seed=29, mode=mode, warps=warps, schedule=schedule, pipe=True)
    orcs = make_oracles(B, W, F, N, seed=29)
    for t0 in range(0, T, 20):
        stats = env.run(20, want_stats=True)
        ref = orc.run_batch(orcs, 20, n_threads=B)
        bad = compare_state(env, orcs, f"pipeline, {warps} warps t={t0 + 20}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.view(np.uint64), ref.view(np.uint64))


def test_out_of_range_tape_entries_are_flagged_not_followed():
    """A foreign tape is untrusted input: an index outside its range is taken as 0 and raises CATS_STATUS_TAPE_RANGE."""
    from r_mabm_b200 import _native as nat
    W, F, N = 200, 20, 5
    env = EmuCats(2, W=W, F=F, N=N, seed=5, mode=0)
    rng = np.random.default_rng(3)
    env.run(3, tape=foreign_tapes(rng, 6, W, F, N, 5, 2, 5).reshape(2, 3, -1))
    assert not (env.status & nat.STATUS_TAPE_RANGE).any()
    off = orc.tape_offsets(W, F, N, 5, 2, 5)
    recs = foreign_tapes(rng, 2, W, F, N, 5, 2, 5).reshape(2, 1, -1)
    gc = recs[0, 0, off["goods_cand"]: off["goods_cand"] + 4 * 5 * (W + F + N)].view(np.int32).reshape(-1, 5)
    gc[:, 0] = 10 ** 6                                    # (a household without a budget draws no candidates: hit them all)
    recs[1, 0, off["job_order"] + 8: off["job_order"] + 12] = np.array([-7], dtype=np.int32).view(np.uint8)
    env.run(1, tape=recs)
    assert (env.status & nat.STATUS_TAPE_RANGE).all()
    assert np.isfinite(env.field("scal").numpy()).all()
