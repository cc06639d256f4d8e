"""Helpers shared by the parity tests: compare a BatchedCats state with CPU-oracle economies."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc

PERSISTED = (["scal", "PA", "perm"] + ["c_" + f for f in orc.C_FIELDS] + ["k_" + f for f in orc.K_FIELDS])
PERSISTED_I32 = ["c_Leff", "k_Leff", "Oc"]
TRANSIENT_F64 = ["c_Ftot", "c_Kdem", "c_Ystock", "c_valinv", "k_Ftot", "k_Ystock"]
TRANSIENT_I32 = ["c_Ld", "c_vac", "k_Ld", "k_vac"]


def make_oracles(B, W, F, N, params=None, seed=0, economy_base=0, bankruptcy_reward=-100.0):
    plist = params if isinstance(params, (list, tuple)) else [params] * B
    return [orc.OracleEconomy(W, F, N, plist[b], seed=seed, economy=economy_base + b,
                              bankruptcy_reward=bankruptcy_reward) for b in range(B)]


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def compare_state(env, oracles, what=""):
    """Bit-exact comparison of every persisted field.  Returns the list of mismatch strings."""
    bad = []
    for name in PERSISTED:
        g = env.field(name).cpu().numpy()
        for b, o in enumerate(oracles):
            ref = o.f64(name)
            if not np.array_equal(bits(g[b]), bits(ref)):
                idx = np.nonzero(bits(g[b]) != bits(ref))[0]
                i = int(idx[0])
                bad.append(f"{what} econ {b} field {name}[{i}] gpu={g[b][i]!r} oracle={ref[i]!r} ({len(idx)} differ)")
                break
    for name in PERSISTED_I32:
        g = env.field(name).cpu().numpy()
        for b, o in enumerate(oracles):
            ref = o.i32(name)
            if not np.array_equal(g[b], ref):
                idx = np.nonzero(g[b] != ref)[0]
                i = int(idx[0])
                bad.append(f"{what} econ {b} field {name}[{i}] gpu={g[b][i]} oracle={ref[i]} ({len(idx)} differ)")
                break
    return bad


def compare_workset(env, debug_img, oracles, what=""):
    """Compare transients from a debug working-set image (after a stop_phase step)."""
    from r_mabm_b200 import _native as nat
    bad = []
    img = debug_img.cpu().numpy()
    for name in TRANSIENT_F64 + TRANSIENT_I32:
        off, cnt, dt = nat.field_info(env.W, env.F, env.N, name, workset=True)
        for b, o in enumerate(oracles):
            raw = img[b, off: off + cnt * (8 if dt == 0 else 4)]
            if dt == 0:
                g, ref = raw.view(np.float64), o.f64(name)
                same = np.array_equal(bits(g), bits(ref))
            else:
                g, ref = raw.view(np.int32), o.i32(name)
                same = np.array_equal(g, ref)
            if not same:
                i = int(np.nonzero(g != ref)[0][0]) if (g != ref).any() else -1
                bad.append(f"{what} econ {b} transient {name}[{i}] gpu={g[i]!r} oracle={ref[i]!r}")
                break
    return bad
