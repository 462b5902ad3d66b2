"""L0: the C oracle's Fr (oracle/fr.h) against Python big ints and the constants of SURVEY.md Appendix A.1."""
import random

import numpy as np

from tests.programs import Q

R = (1 << 256) % Q


def limbs(x):
    return np.array([(x >> (64 * i)) & (2 ** 64 - 1) for i in range(4)], dtype=np.uint64)


def unlimbs(a):
    return sum(int(a[i]) << (64 * i) for i in range(4))


def test_constants(oracle):
    L = oracle.lib()
    one = np.zeros(4, dtype=np.uint64)
    L.orc_fr_from_u64(1, one.ctypes.data)
    assert unlimbs(one) == R == 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
    assert (-pow(Q, -1, 2 ** 64)) % 2 ** 64 == 0xfffffffeffffffff
    assert (R * R) % Q == 0x0748d9d99f59ff1105d314967254398f2b6cedcb87925c23c999e990f3f29c6d
    assert (R * R * R) % Q == 0x6e2a5bb9c8db33e973d13c71c7b5f4181b3e0d188cf06990c62c1807439b73af
    m1 = np.zeros(4, dtype=np.uint64)
    L.orc_fr_neg(one.ctypes.data, m1.ctypes.data)
    assert unlimbs(m1) == 0x5bc8f5f97cd877d899ad88181ce5880ffb38ec08fffb13fcfffffffd00000003


def test_roundtrip_and_ops(oracle):
    rng = random.Random(7)
    vals = [0, 1, 2, Q - 1, Q - 2, 2 ** 64 - 1, 2 ** 64, 2 ** 128, 2 ** 254, R, Q - R] + [rng.randrange(Q) for _ in range(200)]
    m = oracle.from_ints(vals)
    assert [unlimbs(m[i]) for i in range(len(vals))] == [v * R % Q for v in vals]
    assert oracle.to_ints(m) == vals
    L = oracle.lib()
    out = np.zeros(4, dtype=np.uint64)
    for i in range(0, len(vals) - 1):
        a, b = vals[i], vals[i + 1]
        A, B = m[i].copy(), m[i + 1].copy()
        L.orc_fr_mul(A.ctypes.data, B.ctypes.data, out.ctypes.data); assert oracle.to_ints(out[None])[0] == a * b % Q
        L.orc_fr_add(A.ctypes.data, B.ctypes.data, out.ctypes.data); assert oracle.to_ints(out[None])[0] == (a + b) % Q
        L.orc_fr_sub(A.ctypes.data, B.ctypes.data, out.ctypes.data); assert oracle.to_ints(out[None])[0] == (a - b) % Q
        L.orc_fr_neg(A.ctypes.data, out.ctypes.data); assert oracle.to_ints(out[None])[0] == (-a) % Q
        ok = L.orc_fr_invert(A.ctypes.data, out.ctypes.data)
        if a == 0:
            assert ok == 0 and unlimbs(out) == 0
        else:
            assert ok == 1 and oracle.to_ints(out[None])[0] == pow(a, -1, Q)
        assert unlimbs(out) < Q   # fully reduced


def test_pow_and_bits(oracle):
    L = oracle.lib()
    two = oracle.from_ints([2])[0].copy()
    out = np.zeros(4, dtype=np.uint64)
    for e in (0, 1, 63, 64, 65, 128, 200, 254, 255, 256, 300):
        by = np.array([e, 0, 0, 0], dtype=np.uint64)
        L.orc_fr_pow(two.ctypes.data, by.ctypes.data, out.ctypes.data)
        assert oracle.to_ints(out[None])[0] == pow(2, e, Q)
        L.orc_fr_pow_of_2(e, out.ctypes.data)
        assert oracle.to_ints(out[None])[0] == pow(2, e, Q)
    # bits_count KATs of the reference: /root/reference/src/range.rs:198-202
    for v, expect in ((0, 1), (1, 1), (3, 2), (2 ** 128, 129)):
        a = oracle.from_ints([v])[0].copy()
        assert L.orc_bits_count_api(a.ctypes.data) == expect
    # k = bitlen(max-1)+1, with the wrap-around when 2^255 is reduced mod q
    for v in (1, 2, 100, 2 ** 64 - 1, 2 ** 64, 2 ** 252 - 1, 2 ** 253, 2 ** 254 - 1):
        a = oracle.from_ints([v])[0].copy()
        assert L.orc_num_bits_api(a.ctypes.data) == max(1, v.bit_length()) + 1
    a = oracle.from_ints([2 ** 254])[0].copy()
    assert L.orc_num_bits_api(a.ctypes.data) == (2 ** 255 % Q).bit_length()


def test_from_bytes_wide(oracle):
    rng = random.Random(9)
    raw = np.frombuffer(bytes(rng.randrange(256) for _ in range(64 * 50)), dtype=np.uint8).reshape(50, 64)
    m = oracle.from_bytes_wide(raw)
    assert oracle.to_ints(m) == [int.from_bytes(raw[i].tobytes(), "little") % Q for i in range(50)]
