"""Draw tapes written by a generator that is NOT the engine's Philox source (numpy), vectorised over
records -- what a Julia-side recorder on ABCredit.jl would emit (docs/MODEL.md 3.4): outcome-level
permutations, candidate lists and shocks.  Test infrastructure."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc


def _distinct(rng, rows: int, n: int, z: int) -> np.ndarray:
    """[rows, z] int32: z distinct indices out of n per row, in random order."""
    z = min(z, n)
    return np.argsort(rng.random((rows, n)), axis=1)[:, :z].astype(np.int32)


def foreign_tapes(rng, R: int, W: int, F: int, N: int, z_c: int, z_k: int, z_e: int) -> np.ndarray:
    """uint8 [R, tape_bytes]: R independent records."""
    off = orc.tape_offsets(W, F, N, z_c, z_k, z_e)
    H = W + F + N
    out = np.zeros((R, off["bytes"]), dtype=np.uint8)

    def put(name, arr):
        raw = np.ascontiguousarray(arr).reshape(R, -1).view(np.uint8)
        out[:, off[name]: off[name] + raw.shape[1]] = raw

    def cand(rows, n, z):
        c = np.zeros((R, rows, z), dtype=np.int32)        # slots past min(z, n) stay 0: never read
        c[:, :, :min(z, n)] = _distinct(rng, R * rows, n, z).reshape(R, rows, -1)
        return c

    put("u_price_c", rng.random((R, F))); put("u_price_k", rng.random((R, N))); put("u_invest", rng.random((R, F)))
    put("fire_key", rng.integers(0, 2**32, (R, W), dtype=np.uint32))
    put("job_order", rng.permuted(np.tile(np.arange(W, dtype=np.int32), (R, 1)), axis=1))
    put("job_cand", cand(W, F + N, z_e))
    put("cap_order", rng.permuted(np.tile(np.arange(F, dtype=np.int32), (R, 1)), axis=1))
    put("cap_cand", cand(F, N, z_k))
    put("goods_order", rng.permuted(np.tile(np.arange(H, dtype=np.int32), (R, 1)), axis=1))
    put("goods_cand", cand(H, F, z_c))
    return out
