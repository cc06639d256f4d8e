"""CPU: known-answer tests for the oracle's primitives (Philox, deterministic math, sampling)."""
import math

import numpy as np

from oracle import oracle as orc


def _philox(c, k):
    out = np.zeros(4, dtype=np.uint32)
    orc.lib().oracle_philox(*c, *k, out.ctypes.data)
    return [int(x) for x in out]


def test_philox4x32_10_random123_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_det_exp_log_accuracy_and_edges():
    L = orc.lib()
    rng = np.random.default_rng(0)
    for x in np.concatenate([np.linspace(-700, 700, 4001), rng.normal(0, 3, 4000)]):
        assert abs(L.oracle_det_exp(float(x)) - math.exp(x)) <= 4e-16 * math.exp(x)
    for x in np.concatenate([np.exp(np.linspace(-700, 700, 4001)), rng.uniform(0.5, 2, 4000)]):
        ref = math.log(x)
        assert abs(L.oracle_det_log(float(x)) - ref) <= 6e-16 * max(abs(ref), 1e-3)
    assert L.oracle_det_exp(-1e9) == 0.0 and L.oracle_det_exp(float("-inf")) == 0.0
    assert L.oracle_det_exp(1e9) == float("inf")
    assert L.oracle_det_exp(0.0) == 1.0
    assert L.oracle_det_log(0.0) == float("-inf") and math.isnan(L.oracle_det_log(-1.0))
    assert L.oracle_det_log(1.0) == 0.0


def test_det_sum_is_32_lane_strided_butterfly():
    rng = np.random.default_rng(1)
    for n in (0, 1, 5, 20, 32, 33, 100, 1000):
        x = rng.normal(0, 1e3, n)
        lanes = [0.0] * 32
        for lane in range(32):
            a = 0.0
            for i in range(lane, n, 32):
                a = a + x[i]
            lanes[lane] = a
        off = 16
        while off >= 1:
            lanes = [lanes[i] + lanes[i ^ off] for i in range(32)]
            off >>= 1
        got = orc.lib().oracle_det_sum(np.ascontiguousarray(x).ctypes.data, n)
        assert got == lanes[0]
        assert abs(got - math.fsum(x)) <= 1e-9 * max(1.0, abs(math.fsum(x)))


def test_sample_distinct_is_a_partial_permutation():
    rng = np.random.default_rng(2)
    for n, z in [(100, 5), (20, 2), (120, 5), (8, 8), (5, 3), (1000, 8)]:
        for _ in range(200):
            words = rng.integers(0, 2**32, 8, dtype=np.uint32)
            out = np.zeros(8, dtype=np.int32)
            orc.lib().oracle_sample_distinct(words.ctypes.data, n, z, out.ctypes.data)
            sel = out[:z]
            assert len(set(sel.tolist())) == z and sel.min() >= 0 and sel.max() < n
    # first draw is mulhi32(word, n): KAT
    words = np.array([0x80000000, 0, 0, 0, 0, 0, 0, 0], dtype=np.uint32)
    out = np.zeros(8, dtype=np.int32)
    orc.lib().oracle_sample_distinct(words.ctypes.data, 100, 2, out.ctypes.data)
    assert out[0] == 50 and out[1] == 1  # second draw j = 1 + 0 -> position 1 still holds 1


def test_visiting_order_is_a_bijection_and_keyed():
    e = orc.OracleEconomy(1000, 100, 20, seed=9, economy=4)
    perms = []
    for market, n in [(0, 1000), (1, 100), (2, 1120), (2, 7), (2, 1)]:
        out = np.zeros(n, dtype=np.int32)
        orc.lib().oracle_perm(e._h, market, n, out.ctypes.data)
        assert sorted(out.tolist()) == list(range(n))
        perms.append(out)
    e2 = orc.OracleEconomy(1000, 100, 20, seed=9, economy=5)
    out2 = np.zeros(1120, dtype=np.int32)
    orc.lib().oracle_perm(e2._h, 2, 1120, out2.ctypes.data)
    assert not np.array_equal(out2, perms[2])  # another economy, another order
    # not the identity, and roughly uniform displacement
    assert np.mean(perms[2] == np.arange(1120)) < 0.02
