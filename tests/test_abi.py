"""CPU: libcats_b200.so loads without a GPU and exports every symbol include/*.h declares;
the layout queries (no compute) agree with the oracle's tape layout."""
import ctypes as C
import re
from pathlib import Path

import pytest

from r_mabm_b200 import _native as nat

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cats_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = nat.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/*.h but not exported"
    assert sorted(nat.EXPORTS) == syms


def test_layout_queries():
    lib = nat.load()
    assert lib.cats_abi_version() == nat.ABI_VERSION == 2
    blob = int(lib.cats_blob_bytes(1000, 100, 20))
    assert blob % 128 == 0 and 35000 < blob < 45000
    names = nat.field_names()
    assert {"scal", "c_P", "c_Q", "k_P", "c_Leff", "k_Leff", "PA", "perm", "Oc"} <= set(names)
    owner = {}   # every byte of the blob belongs to at most one field element
    for n in names:
        off, cnt, dt = nat.field_info(1000, 100, 20, n)
        esz, st = (8 if dt == 0 else 4), nat.field_stride(n)
        assert off % esz == 0 and off + ((cnt - 1) * st + 1) * esz <= blob
        assert st == (2 if n in ("PA", "perm") else 1)          # households are {PA, perm} records
        for i in range(cnt):
            for byte in range(off + i * st * esz, off + i * st * esz + esz, 4):
                assert byte not in owner, f"{n}[{i}] overlaps {owner[byte]}"
                owner[byte] = n
    with pytest.raises(ValueError):
        nat.field_stride("no_such_field")
    assert nat.field_info(1000, 100, 20, "c_P")[1] == 100 and nat.field_info(1000, 100, 20, "PA")[1] == 1120
    with pytest.raises(ValueError):
        nat.field_info(1000, 100, 20, "no_such_field")
    with pytest.raises(ValueError):
        nat.field_info(0, 100, 20, "c_P")
    assert int(lib.cats_workset_bytes(1000, 100, 20)) > 0
    assert nat.field_info(1000, 100, 20, "c_vac", workset=True)[2] == 1


def test_tape_layout_matches_the_oracle(oracle_mod):
    lib = nat.load()
    for (W, F, N, zc, zk, ze) in [(1000, 100, 20, 5, 2, 5), (120, 12, 4, 4, 2, 3), (10000, 1000, 200, 5, 2, 5)]:
        out = (C.c_uint64 * 11)()
        assert lib.cats_tape_offsets(W, F, N, zc, zk, ze, out) == 0
        ref = oracle_mod.tape_offsets(W, F, N, zc, zk, ze)
        assert [int(v) for v in out] == list(ref.values())
        assert int(lib.cats_tape_bytes(W, F, N, zc, zk, ze)) == ref["bytes"]


def test_argument_errors_without_a_gpu():
    lib = nat.load()
    a = nat.StepArgs()
    a.W, a.F, a.N, a.B, a.n_periods = 1000, 100, 20, 1, 1
    assert lib.cats_step(C.byref(a)) == nat.EINVAL            # null blobs
    assert b"null" in lib.cats_last_error()
    assert lib.cats_init(None, 1, 1000, 100, 20, None, 0, None) == nat.EINVAL
    assert lib.cats_reset_masked(None, 1, 1000, 100, 20, None, 0, None, None) == nat.EINVAL
    assert int(lib.cats_host_scratch_bytes(16, 2, 1)) >= 16 * (2 * 16 + 2 + 2 * 16 + 2 * 8 + 1 + 4 + 128)
