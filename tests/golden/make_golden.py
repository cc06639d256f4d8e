"""Regenerates the golden fixtures from the CPU oracle:  python tests/golden/make_golden.py

NOT reference vectors: the reference (ABCredit.jl through juliacall) cannot run in this environment
and its own tests hold no numeric fixtures for the step ("parity unpinned").  These files pin the
in-repo model spec (docs/MODEL.md) so that an accidental change of either implementation shows up:
  small_snapshots.npz  -- full state of a (W=120, F=12, N=4) economy, seeds 0..3, after 1/10/100 periods
  default_digests.json -- sha256 of the full state of default-size economies, seeds 0..3, after 1/10/100
  tape_record.npz      -- one draw-tape record (seed 3, economy 5, period 7) of the default-size economy
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

HERE = Path(__file__).resolve().parent
PERIODS = (1, 10, 100)
FIELDS_F64 = ["scal", "PA", "perm"] + ["c_" + f for f in orc.C_FIELDS] + ["k_" + f for f in orc.K_FIELDS]
FIELDS_I32 = ["c_Leff", "k_Leff", "Oc"]


def state_bytes(e) -> bytes:
    return b"".join(e.f64(n).tobytes() for n in FIELDS_F64) + b"".join(e.i32(n).tobytes() for n in FIELDS_I32)


def main():
    small = {}
    for seed in range(4):
        e = orc.OracleEconomy(120, 12, 4, seed=seed, economy=seed)
        t = 0
        for p in PERIODS:
            while t < p:
                e.step(); t += 1
            for n in FIELDS_F64:
                small[f"s{seed}_t{p}_{n}"] = e.f64(n).copy()
            for n in FIELDS_I32:
                small[f"s{seed}_t{p}_{n}"] = e.i32(n).copy()
    np.savez_compressed(HERE / "small_snapshots.npz", **small)

    digests = {}
    for seed in range(4):
        e = orc.OracleEconomy(1000, 100, 20, seed=seed, economy=100 + seed)
        t = 0
        for p in PERIODS:
            while t < p:
                e.step(); t += 1
            digests[f"seed{seed}_t{p}"] = {
                "sha256": hashlib.sha256(state_bytes(e)).hexdigest(),
                "Y_real": e.scal("Y_real"), "Un": e.scal("Un"), "price": e.scal("price"),
            }
    (HERE / "default_digests.json").write_text(json.dumps(digests, indent=1))

    e = orc.OracleEconomy(1000, 100, 20, seed=3, economy=5)
    for _ in range(7):
        e.step()
    np.savez_compressed(HERE / "tape_record.npz", record=e.fill_tape())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
