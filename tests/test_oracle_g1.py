"""G1 / commitment oracle (second half): oracle/g1.c (Montgomery Fp, Jacobian G1, the crate's serial
Pippenger) against the affine big-int model of oracle/pymodel.py, and the constants against the curve's defining identities."""
import random

import numpy as np

from tests.programs import Q


def test_curve_constants(oracle):
    """p and q from the BLS parameter x = -0xd201000000010000; generator on y^2 = x^3 + 4 and of order q."""
    from oracle import pymodel as pm
    x = -0xd201000000010000
    assert Q == x ** 4 - x ** 2 + 1
    assert pm.P_FIELD == (x - 1) ** 2 * (x ** 4 - x ** 2 + 1) // 3 + x == oracle.P_FIELD
    G = pm.G1_GENERATOR
    assert pm.g1_on_curve(G) and pm.g1_mul(Q, G) is None and pm.g1_mul(Q - 1, G) == pm.g1_neg(G)
    assert oracle.g1_to_ints(oracle.g1_generator()) == [G]


def test_group_law(oracle):
    from oracle import pymodel as pm
    G = pm.G1_GENERATOR
    rnd = random.Random(3)
    ks = [0, 1, 2, 3, Q - 1, Q - 2] + [rnd.randrange(Q) for _ in range(6)]
    pts = oracle.g1_mul(np.repeat(oracle.g1_generator(), len(ks), axis=0), oracle.from_ints(ks))
    exp = [pm.g1_mul(k, G) for k in ks]
    assert oracle.g1_to_ints(pts) == exp
    lhs = [exp[2], exp[2], exp[2], None, exp[6], None]
    rhs = [exp[3], exp[2], pm.g1_neg(exp[2]), exp[7], None, None]          # generic, doubling, inverse, infinities
    assert oracle.g1_to_ints(oracle.g1_add(oracle.g1_from_ints(lhs), oracle.g1_from_ints(rhs))) == [pm.g1_add(a, b) for a, b in zip(lhs, rhs)]


def test_msm_and_srs(oracle):
    from oracle import pymodel as pm
    G = pm.G1_GENERATOR
    rnd = random.Random(4)
    for n in (1, 5, 31, 32, 40, 72):
        sc = [rnd.randrange(Q) for _ in range(n)]
        sc[0] = 1
        if n > 3:
            sc[2] = 0; sc[3] = 1
        base = [pm.g1_mul(rnd.randrange(Q), G) for _ in range(n)]
        if n > 5:
            base[4] = None; base[5] = base[1]                               # a point at infinity and a repeated point
        got = oracle.g1_to_ints(oracle.g1_msm(oracle.g1_from_ints(base), oracle.from_ints(sc)))[0]
        assert got == pm.g1_msm(sc, base), n
    assert oracle.g1_to_ints(oracle.srs_powers(oracle.from_ints([12345]), 6)) == pm.srs_powers(12345, 6)


def _pts(lst):
    return [None if p is None else (int(p[0], 16), int(p[1], 16)) for p in lst]


def test_g1_golden(oracle, golden, golden_g1):
    """oracle/g1.c against the committed fixtures tests/golden/g1.json (generated from the big-int model by oracle/gen_golden.py)."""
    from tests.programs import run_oracle
    from oracle import pymodel as pm
    g = oracle.g1_generator()
    ks = [int(k, 16) for k in golden_g1["multiples"]]
    for k in ks[:2] + ks[-2:]:                                                   # the fixture is what the big-int model produces
        assert _pts([golden_g1["multiples"][hex(k)]]) == [pm.g1_mul(k, pm.G1_GENERATOR)]
    got = oracle.g1_to_ints(oracle.g1_mul(np.repeat(g, len(ks), axis=0), oracle.from_ints(ks)))
    assert got == _pts(golden_g1["multiples"].values())
    beta = int(golden_g1["srs"]["beta"], 16)
    assert oracle.g1_to_ints(oracle.srs_powers(oracle.from_ints([beta]), 8)) == _pts(golden_g1["srs"]["powers"])
    for n, vec in golden_g1["msm"].items():
        pts = oracle.g1_from_ints(_pts(vec["points"]))
        sc = oracle.from_ints([int(v, 16) for v in vec["scalars"]])
        assert oracle.g1_to_ints(oracle.g1_msm(pts, sc)) == _pts([vec["sum"]]), n
    for name, exp in golden_g1["wire_commitments"].items():
        _s, oc = run_oracle(golden[name]["program"], return_composer=True)
        polys = oc.wire_polynomials()
        srs = oracle.srs_powers(oracle.from_ints([beta]), 1 << exp["log_n"])
        assert [oracle.g1_to_ints(oracle.g1_msm(srs, polys[w]))[0] for w in range(4)] == _pts(exp["commitments"]), name
