"""Replay mode on the GPU at width (north_star: "a replay mode that injects the reference's own random
draws"; BASELINE.json config 2) and the scaled economy in its steady state.

* foreign tapes: numpy-generated outcome-level draws (tape_util.foreign_tapes) -- orders that are not
  Feistel permutations, candidate lists that are not Philox-derived -- through the MODE 0 kernel at
  1,024 economies x 100 periods, streamed in chunks, against oracles replaying the same records;
* the oracle's own Philox record for every one of 1,024 economies: tape replay == Philox mode == oracle;
* default size with tape z != (5, 2, 5) and with more than 128 job seekers in a period;
* the 10x economy (config 5) for 300 burn-in + 20 periods, 4 economies.

Oracle = the in-repo CPU restatement (parity unpinned w.r.t. ABCredit.jl).  Bar: bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from parity_util import compare_state, make_oracles
from tape_util import foreign_tapes

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 4


def _env(B, **kw):
    from r_mabm_b200.engine import BatchedCats
    return BatchedCats(B, **kw)


def test_config2_foreign_tape_replay_at_width():
    """1,024 default-size economies x 100 periods driven by foreign tapes: 8 distinct tape streams, each
    replayed by 128 economies spread over the launch.  The 8 streams equal oracles replaying the same
    records (state + every period's statistics), and all copies of a stream are identical."""
    B, K, T, CH = 1024, 8, 100, 10
    W, F, N = 1000, 100, 20
    env = _env(B, seed=1)
    orcs = make_oracles(K, W, F, N, seed=99)           # seed and id are irrelevant under a tape
    rng = np.random.default_rng(2026)
    idx = torch.arange(B, device="cuda") % K            # economy b replays stream b % K
    for t0 in range(0, T, CH):
        recs = foreign_tapes(rng, K * CH, W, F, N, 5, 2, 5).reshape(K, CH, -1)
        tape = torch.from_numpy(recs).cuda()[idx].contiguous()       # [B, CH, bytes]
        stats = env.run(CH, want_stats=True, tape=tape.view(-1))
        ref = orc.run_batch(orcs, CH, n_threads=min(THREADS, K), tape=recs, fill_tape=False)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"t={t0 + CH}")   # compares economies 0..K-1
        assert bad == [], "\n".join(bad)
        g = stats.cpu().numpy()
        assert np.array_equal(g[:K].view(np.uint64), ref.view(np.uint64))
        blobs = env.blobs.view(B // K, K, -1)
        assert torch.equal(blobs, blobs[:1].expand_as(blobs)), f"copies differ at t={t0 + CH}"
        assert np.array_equal(g.reshape(B // K, K, CH, -1).view(np.uint64),
                              np.broadcast_to(g[:K].view(np.uint64), (B // K, K, CH, g.shape[-1])))
    assert 0.0 < float(env.scalar("Un").mean()) < 0.5


def test_config2_philox_record_replay_equals_philox_mode_at_width():
    """Every one of 1,024 economies replays the record of its own Philox stream (written by the oracle):
    the tape kernel (MODE 0), the production kernel (MODE 2) and the oracles agree on every bit."""
    B, T, CH = 1024, 100, 10
    W, F, N = 1000, 100, 20
    a, b = _env(B, seed=7), _env(B, seed=7)
    orcs = make_oracles(B, W, F, N, seed=7)
    tb = a.tape_bytes()
    for t0 in range(0, T, CH):
        recs = np.zeros((B, CH, tb), dtype=np.uint8)
        ref = orc.run_batch(orcs, CH, n_threads=THREADS, tape=recs, fill_tape=True)
        sa = a.run(CH, want_stats=True, tape=torch.from_numpy(recs).cuda().view(-1))
        sb = b.run(CH, want_stats=True)
        torch.cuda.synchronize()
        assert torch.equal(a.blobs, b.blobs), t0
        assert torch.equal(sa, sb)
        assert np.array_equal(sa.cpu().numpy().view(np.uint64), ref.view(np.uint64))
    sample = [0, 1, 137, 255, 256, 511, 777, 1023]
    sub = _Sub(a, sample)
    bad = compare_state(sub, [orcs[i] for i in sample], f"t={T}")
    assert bad == [], "\n".join(bad)


class _Sub:
    """A view of some economies of an engine, for compare_state."""

    def __init__(self, env, ids):
        self.env, self.ids = env, torch.tensor(ids, device="cuda")
        self.W, self.F, self.N = env.W, env.F, env.N

    def field(self, name):
        return self.env.field(name)[self.ids]


@pytest.mark.parametrize("zs", [(5, 2, 5), (4, 3, 6), (8, 8, 8)])
def test_foreign_tape_default_size(zs):
    """Foreign tapes at the default size for 40 periods from a cold start (period 1 has 1,000 job seekers:
    the path for more than 128 seekers and tape orders), also with tape z != the defaults."""
    B, T = 4, 40
    W, F, N = 1000, 100, 20
    p = dict(orc.DEFAULT_PARAMS, z_c=zs[0], z_k=zs[1], z_e=zs[2])
    env = _env(B, params=dict(p), seed=3)
    orcs = make_oracles(B, W, F, N, params=p, seed=3)
    rng = np.random.default_rng(zs[0] * 100 + zs[1] * 10 + zs[2])
    for t0 in range(0, T, 8):
        recs = foreign_tapes(rng, B * 8, W, F, N, *zs).reshape(B, 8, -1)
        assert recs.shape[2] == env.tape_bytes()
        stats = env.run(8, want_stats=True, tape=torch.from_numpy(recs).cuda().view(-1))
        ref = orc.run_batch(orcs, 8, n_threads=B, tape=recs, fill_tape=False)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"z={zs} t={t0 + 8}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))


def test_foreign_tape_single_period_launches_with_actions():
    """A tape with RL actions on top (one period per launch): the override happens after the taped
    heuristic draws, as in cats.py:363-370."""
    B, A, W, F, N = 3, 2, 400, 40, 8
    env = _env(B, W=W, F=F, N=N, n_agents=A, seed=5)
    orcs = make_oracles(B, W, F, N, seed=5)
    rng = np.random.default_rng(11)
    for t in range(25):
        recs = foreign_tapes(rng, B, W, F, N, 5, 2, 5)
        acts = rng.uniform(0.8, 1.2, size=(B, A, 2))
        obs, rew, term, _ = env.step(torch.from_numpy(acts).cuda(), tape=torch.from_numpy(recs).cuda().view(-1))
        torch.cuda.synchronize()
        for b, o in enumerate(orcs):
            r_obs, r_rew = o.step(actions=acts[b], flags=0, tape=recs[b])
            assert np.array_equal(obs[b].cpu().numpy().view(np.uint64), r_obs.view(np.uint64)), (t, b)
            assert np.array_equal(rew[b].cpu().numpy().view(np.uint64), r_rew.view(np.uint64)), (t, b)
        assert compare_state(env, orcs, f"t={t}") == []


def test_scaled_economy_10x_steady_state():
    """BASELINE.json config 5 after the reference's burn-in: 4 economies of W=10,000, F=1,000, N=200 for
    300 + 20 periods (the steady state: few seekers, sell-outs, bankruptcies), state compared every 20
    periods during the burn-in and every period after it."""
    B, W, F, N = 4, 10000, 1000, 200
    env = _env(B, W=W, F=F, N=N, seed=77)
    orcs = make_oracles(B, W, F, N, seed=77)
    for t0 in range(0, 300, 20):
        stats = env.run(20, want_stats=True)
        ref = orc.run_batch(orcs, 20, n_threads=B)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"10x burn-in t={t0 + 20}")
        assert bad == [], "\n".join(bad)
        assert np.array_equal(stats.cpu().numpy().view(np.uint64), ref.view(np.uint64))
    for t in range(20):
        env.run(1)
        orc.run_batch(orcs, 1, n_threads=B, want_stats=False)
        torch.cuda.synchronize()
        bad = compare_state(env, orcs, f"10x t={301 + t}")
        assert bad == [], "\n".join(bad)
    un = env.scalar("Un").cpu().numpy()
    assert (un < 0.3).all()


def test_active_mask_with_tape_or_debug_stop_is_rejected():
    """the masked kernel is a production instantiation; combining `active` with a tape, a debug
    stop or a debug image must fail instead of silently drawing from Philox."""
    env = _env(2, W=100, F=10, N=3, seed=1)
    act = torch.ones(2, dtype=torch.uint8, device="cuda")
    tape = torch.zeros(2 * env.tape_bytes(), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        env.run(1, tape=tape, active=act)
    with pytest.raises(ValueError):
        env.step(burnin=True, stop_phase=5, active=act)
    import ctypes as C
    from r_mabm_b200 import _native as nat
    a = env._args(1, env.flags | nat.FLAG_BURNIN, want_io=False)
    a.d_active = act.data_ptr(); a.d_tape = tape.data_ptr()
    a.tape_z[0], a.tape_z[1], a.tape_z[2] = 5, 2, 5
    assert env._lib.cats_step(C.byref(a)) == nat.EINVAL


def test_out_of_range_tape_entries_are_flagged_not_followed():
    """A foreign tape is untrusted input: an index outside its range is taken as 0 and raises CATS_STATUS_TAPE_RANGE
    (a wild shared-memory access otherwise)."""
    from r_mabm_b200 import _native as nat
    W, F, N = 1000, 100, 20
    env = _env(2, seed=5)
    env.run(20)
    rng = np.random.default_rng(3)
    off = orc.tape_offsets(W, F, N, 5, 2, 5)
    recs = foreign_tapes(rng, 2, W, F, N, 5, 2, 5).reshape(2, 1, -1)
    env.run(1, tape=torch.from_numpy(recs).cuda().view(-1))
    torch.cuda.synchronize()
    assert int(env._status.max()) == 0
    recs = foreign_tapes(rng, 2, W, F, N, 5, 2, 5).reshape(2, 1, -1)
    gc = recs[0, 0, off["goods_cand"]: off["goods_cand"] + 4 * 5 * (W + F + N)].view(np.int32).reshape(-1, 5)
    gc[:, 0] = 10 ** 6                                    # (a household without a budget draws no candidates: hit them all)
    recs[1, 0, off["job_order"] + 8: off["job_order"] + 12] = np.array([-7], dtype=np.int32).view(np.uint8)
    env.run(1, tape=torch.from_numpy(recs).cuda().view(-1))
    torch.cuda.synchronize()
    assert (env._status.cpu().numpy() & nat.STATUS_TAPE_RANGE).all()
    assert torch.isfinite(env.field("scal")).all()
