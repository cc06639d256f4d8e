"""CATS_FLAG_STRICT on the GPU: the plain arithmetic of docs/MODEL.md 4.6 (b / P, demand summed in
visiting order in fp64, no reciprocals) against the oracle in the same mode, bit for bit -- Philox
draws (MODE 1 and the default-configuration MODE 2 kernel), foreign tapes (MODE 0), phase by phase
through the markets, every block shape.  The fast (default) rules are this repository's own choice
and diverge from the plain ones in integer state within a few periods; the test pins that too, so
nobody mistakes one mode for the other.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import compare_state, compare_workset, make_oracles
from tape_util import foreign_tapes

pytestmark = pytest.mark.gpu
STRICT = orc.FLAG_STRICT


def _env(B, **kw):
    from r_mabm_b200.engine import BatchedCats
    return BatchedCats(B, strict=True, **kw)


@pytest.mark.parametrize("W,F,N,zs", [(1000, 100, 20, (5, 2, 5)), (1000, 100, 20, (4, 3, 6)), (300, 30, 8, (5, 2, 5)),
                                       (130, 7, 2, (8, 8, 8))])
def test_strict_mode_matches_strict_oracle(W, F, N, zs):
    B, T = 6, 60
    p = dict(orc.DEFAULT_PARAMS, z_c=zs[0], z_k=zs[1], z_e=zs[2])
    env = _env(B, W=W, F=F, N=N, params=dict(p), seed=13)
    orcs = make_oracles(B, W, F, N, params=p, seed=13)
    for t0 in range(0, T, 10):
        stats = env.run(10, want_stats=True)
        ref = orc.run_batch(orcs, 10, n_threads=B, flags=STRICT)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"strict {(W, F, N, zs)} t={t0 + 10}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))


@pytest.mark.parametrize("phase", [16, 18, 19, 20, 22])
def test_strict_phase_by_phase(phase):
    B = 3
    env = _env(B, seed=11)
    orcs = make_oracles(B, 1000, 100, 20, seed=11)
    env.run(3)
    orc.run_batch(orcs, 3, n_threads=B, want_stats=False, flags=STRICT)
    dbg = torch.zeros((B, env.workset_bytes()), dtype=torch.uint8, device="cuda")
    env.step(burnin=True, stop_phase=phase, debug=dbg)
    for o in orcs:
        o.set_stop_phase(phase); o.step(flags=orc.FLAG_BURNIN | STRICT)
    torch.cuda.synchronize()
    bad = compare_state(env, orcs, f"strict phase {phase}")
    if phase < 21:
        bad += compare_workset(env, dbg, orcs, f"strict phase {phase}")
    assert bad == [], "\n".join(bad)


def test_strict_foreign_tape_replay():
    """The combination a pin against ABCredit.jl would use: foreign draws + plain arithmetic."""
    B, W, F, N = 4, 1000, 100, 20
    env = _env(B, seed=3)
    orcs = make_oracles(B, W, F, N, seed=3)
    rng = np.random.default_rng(8)
    for t0 in range(0, 40, 8):
        recs = foreign_tapes(rng, B * 8, W, F, N, 5, 2, 5).reshape(B, 8, -1)
        stats = env.run(8, want_stats=True, tape=torch.from_numpy(recs).cuda().view(-1))
        ref = orc.run_batch(orcs, 8, n_threads=B, tape=recs, fill_tape=False, flags=STRICT)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"strict tape t={t0 + 8}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))


def test_strict_rl_step_and_block_shapes():
    """RL actions, rewards and observations in strict mode; every block shape gives the same bits."""
    from r_mabm_b200 import _native as nat
    from r_mabm_b200.engine import BatchedCats
    lib = nat.load()
    B, A, W, F, N = 29, 2, 200, 20, 5
    rng = np.random.default_rng(4)
    acts = rng.uniform(0.8, 1.2, size=(B, A, 2))
    orcs = make_oracles(B, W, F, N, seed=21)
    orc.run_batch(orcs, 7, n_threads=8, want_stats=False, flags=STRICT)
    ref_out = [o.step(actions=acts[b], flags=STRICT) for b, o in enumerate(orcs)]
    ref = None
    try:
        for g in (0, 1, 2, 4, 8, 12, 14):
            lib.cats_debug_set_group(g)
            env = BatchedCats(B, W=W, F=F, N=N, n_agents=A, seed=21, strict=True)
            env.run(7)
            obs, rew, term, _ = env.step(torch.from_numpy(acts).cuda())
            torch.cuda.synchronize()
            if ref is None:
                assert compare_state(env, orcs, "strict rl") == []
                for b in range(B):
                    assert np.array_equal(obs[b].cpu().numpy().view(np.uint64), ref_out[b][0].view(np.uint64))
                    assert np.array_equal(rew[b].cpu().numpy().view(np.uint64), ref_out[b][1].view(np.uint64))
                ref = env.blobs.clone()
            else:
                assert torch.equal(env.blobs, ref), f"group {g}"
    finally:
        lib.cats_debug_set_group(0)


def test_fast_and_strict_rules_part_ways_within_a_few_periods():
    """MODEL.md 4.6 deviations register: the default (fast) rules are NOT inside the 1e-9 / 100-period band
    of the plain ones -- employment links differ within a few dozen periods."""
    B = 4
    fast = __import__("r_mabm_b200.engine", fromlist=["BatchedCats"]).BatchedCats(B, seed=0)
    strict = _env(B, seed=0)
    first = [None] * B
    for t in range(1, 61):
        fast.run(1); strict.run(1)
        diff = (fast.field("Oc") != strict.field("Oc")).any(dim=1).cpu().numpy()
        for b in range(B):
            if first[b] is None and diff[b]:
                first[b] = t
    assert all(f is not None and f <= 60 for f in first), first


def test_strict_with_active_mask_is_rejected():
    env = _env(2, W=100, F=10, N=3, seed=1)
    with pytest.raises(ValueError):
        env.run(1, active=torch.ones(2, dtype=torch.uint8, device="cuda"))
