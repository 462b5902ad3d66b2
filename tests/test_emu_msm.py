"""CPU tests of the commitment path (pg_msm, pg_srs_powers, pg_commit_wire_polynomials, pg_g1_op) through tests/emu: the
Fp / G1 formulas (g1.cuh) and the bucket-method bodies (msm.cuh) run on the loop backend and are compared with the oracle.
The CUDA launches (and the CUB sort) are checked on the B200 by tests/test_gpu_parity.py."""
import random

import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from tests.engine_runner import run_engine
from tests.programs import Q, run_oracle, synth_wide
from tests.test_emu_engine import emu  # noqa: F401  (fixture)


def to_engine(points13: np.ndarray) -> np.ndarray:
    """oracle points (n, 13) -> engine points (n, 12): infinity becomes the all-zero pair."""
    p = np.ascontiguousarray(points13, dtype=np.uint64).reshape(-1, 13)
    out = p[:, :12].copy()
    out[(p[:, 12] & np.uint64(0xffffffff)) != 0] = 0
    return out


def to_oracle(points12: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(points12, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((p.shape[0], 13), dtype=np.uint64)
    out[:, :12] = p
    out[:, 12] = (~p.any(axis=1)).astype(np.uint64)
    return out


def test_emu_group_law(emu, oracle):
    c = pg.StandardComposer(_cdll=emu)
    g = oracle.g1_generator()
    rnd = random.Random(8)
    ks = [1, 2, 3, Q - 1, Q - 2, rnd.randrange(Q), rnd.randrange(Q)]
    pts = oracle.g1_mul(np.repeat(g, len(ks), axis=0), oracle.from_ints(ks))
    inf = np.zeros((1, 13), dtype=np.uint64); inf[0, 12] = 1
    lhs = np.concatenate([pts[0:1], pts[1:2], pts[1:2], inf, pts[5:6], inf, pts[2:3]])
    rhs = np.concatenate([pts[1:2], pts[1:2], pts[4:5], pts[6:7], inf, inf, pts[3:4]])   # generic, doubling, inverse (2G + (q-2)G), inf+P, P+inf, inf+inf, 3G+(q-1)G
    got = c.g1_op(0, to_engine(lhs), to_engine(rhs))
    assert np.array_equal(to_oracle(got), oracle.g1_add(lhs, rhs))
    flags = c.g1_op(1, to_engine(np.concatenate([pts, inf])))
    assert flags[:, 0].tolist() == [1] * (len(ks) + 1)
    bad = to_engine(pts[:1]); bad[0, 0] ^= np.uint64(1)
    assert c.g1_op(1, bad)[0, 0] == 0
    # fixed-base multiples against the oracle, including 0 and q-1
    sc = oracle.from_ints([0, 1, 5, Q - 1, rnd.randrange(Q)])
    assert np.array_equal(to_oracle(c.g1_fixed_base_mul(sc)), oracle.g1_mul(np.repeat(g, 5, axis=0), sc))


def test_emu_windowed_fixed_base(emu, oracle, monkeypatch):
    """From 512 scalars on pg_g1_fixed_base_mul / pg_srs_powers go through the table of window multiples and the batch
    normalisation (msm.cuh: G1WindowTableBody, G1FixedBaseWindowedBody, G1BatchAffineBody) and the device-side powers of beta."""
    monkeypatch.setenv("PG_FB_MIN", "512")                   # (the table pays off from 2^14 scalars on; forced here at test sizes)
    c = pg.StandardComposer(_cdll=emu)
    rnd = random.Random(21)
    n = 530                                                  # not a multiple of the normalisation chunk
    ks = [0, 1, 255, 256, 2 ** 248, Q - 1, 2 ** 64 - 1] + [rnd.randrange(Q) for _ in range(n - 7)]
    ks[40] = 0; ks[n - 1] = 0                                # points at infinity inside and at the end of a chunk
    sc = oracle.from_ints(ks)
    g = oracle.g1_generator()
    assert np.array_equal(to_oracle(c.g1_fixed_base_mul(sc)), oracle.g1_mul(np.repeat(g, n, axis=0), sc))
    base = oracle.g1_mul(g, oracle.from_ints([rnd.randrange(Q)]))
    assert np.array_equal(to_oracle(c.g1_fixed_base_mul(sc[:520], base=to_engine(base))), oracle.g1_mul(np.repeat(base, 520, axis=0), sc[:520]))
    beta = oracle.from_ints([rnd.randrange(Q)])
    assert np.array_equal(to_oracle(c.srs_powers(beta[0], 600)), oracle.srs_powers(beta, 600))


@pytest.mark.parametrize("n", [1, 2, 7, 33, 100, 300])
def test_emu_msm_matches_oracle(emu, oracle, n):
    c = pg.StandardComposer(_cdll=emu)
    rnd = random.Random(n)
    srs = oracle.srs_powers(oracle.from_ints([rnd.randrange(Q)]), n)
    sc = synth_wide(40 + n, n)
    sc[0] = 1
    if n > 4:
        sc[2] = 0; sc[3] = 1; sc[4] = Q - 1
    if n > 40:
        srs[7] = srs[6]                                         # a repeated point: equal-point additions inside a bucket
        sc[7] = sc[6]
        srs[9, :12] = 0; srs[9, 12] = 1                         # a point at infinity
    got = c.msm(to_engine(srs), oracle.from_ints(sc))
    assert np.array_equal(to_oracle(got.reshape(1, 12)), oracle.g1_msm(srs, oracle.from_ints(sc)))


def test_emu_msm_edge_cases(emu, oracle):
    c = pg.StandardComposer(_cdll=emu)
    srs = oracle.srs_powers(oracle.from_ints([77]), 16)
    assert not c.msm(to_engine(srs), oracle.from_ints([0] * 16)).any()                  # all-zero scalars: infinity
    assert not c.msm(np.zeros((0, 12), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64)).any()   # empty sum
    same = np.repeat(srs[3:4], 16, axis=0)                                              # 16 copies of one point, scalar 1 each: 16 * P
    got = c.msm(to_engine(same), oracle.from_ints([1] * 16))
    assert np.array_equal(to_oracle(got.reshape(1, 12)), oracle.g1_mul(srs[3:4], oracle.from_ints([16])))
    pair = np.concatenate([srs[3:4], srs[3:4]])                                         # P - P
    assert not c.msm(to_engine(pair), oracle.from_ints([5, Q - 5])).any()


def test_emu_srs_and_wire_commitments(emu, oracle, golden):
    spec = golden["kat_range_check_0_ok"]
    _s, oc = run_oracle(spec["program"], return_composer=True)
    _snap, c = run_engine(spec["program"], lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
    k = c.domain_log_size()
    beta = oracle.from_ints([0x1234567890abcdef1234567890abcdef])
    srs = c.srs_powers(beta[0], 1 << k)
    assert np.array_equal(to_oracle(srs), oracle.srs_powers(beta, 1 << k))
    got = c.commit_wire_polynomials(srs)
    polys = oc.wire_polynomials()
    for w in range(4):
        assert np.array_equal(to_oracle(got[w:w + 1]), oracle.g1_msm(to_oracle(srs), polys[w])), w


def check_g1_golden(make_composer, oracle, golden, golden_g1):
    """Engine (emu or CUDA) against tests/golden/g1.json: generator multiples, SRS powers, sums, wire commitments."""
    from tests.test_oracle_g1 import _pts
    c = make_composer()
    ks = [int(k, 16) for k in golden_g1["multiples"]]
    assert oracle.g1_to_ints(to_oracle(c.g1_fixed_base_mul(oracle.from_ints(ks)))) == _pts(golden_g1["multiples"].values())
    beta = oracle.from_ints([int(golden_g1["srs"]["beta"], 16)])
    assert oracle.g1_to_ints(to_oracle(c.srs_powers(beta[0], 8))) == _pts(golden_g1["srs"]["powers"])
    for n, vec in golden_g1["msm"].items():
        pts = to_engine(oracle.g1_from_ints(_pts(vec["points"])))
        got = c.msm(pts, oracle.from_ints([int(v, 16) for v in vec["scalars"]]))
        assert oracle.g1_to_ints(to_oracle(got.reshape(1, 12))) == _pts([vec["sum"]]), n
    for name, exp in golden_g1["wire_commitments"].items():
        _snap, cc = run_engine(golden[name]["program"], make_composer, oracle, return_composer=True)
        assert cc.domain_log_size() == exp["log_n"]
        got = cc.commit_wire_polynomials(cc.srs_powers(beta[0], 1 << exp["log_n"]))
        assert oracle.g1_to_ints(to_oracle(got)) == _pts(exp["commitments"]), name


def test_emu_g1_golden(emu, oracle, golden, golden_g1):
    check_g1_golden(lambda: pg.StandardComposer(_cdll=emu), oracle, golden, golden_g1)


def test_emu_argument_errors(emu, oracle, golden):
    """Bad arguments of the evaluation-domain / commitment entry points are rejected with PG_ERR_ARG, not executed."""
    import ctypes as C
    spec = golden["kat_range_check_0_ok"]
    _snap, c = run_engine(spec["program"], lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
    k = c.domain_log_size()
    srs = c.srs_powers(oracle.from_ints([9])[0], 1 << k)
    with pytest.raises(pg.EngineError) as e:
        c.commit_wire_polynomials(srs[: (1 << k) - 1])                       # SRS one power short: polynomial degree too large
    assert e.value.code == -2
    with pytest.raises(pg.EngineError):
        c.commit_wire_polynomials(srs, log_n=k - 1)                          # domain smaller than the circuit
    buf = np.zeros((2, 4), dtype=np.uint64)
    p = buf.ctypes.data_as(C.c_void_p)
    assert c._L.pg_fft(c._ctx, 33, 0, p, p, 0) == -2                         # beyond the two-adicity of Fr
    assert c._L.pg_fft(c._ctx, 1, 0, None, p, 0) == -2
    assert c._L.pg_msm(c._ctx, 2, None, p, p, 0) == -2
    assert c._L.pg_wire_polynomials(c._ctx, k, None, 0) == -2
    with pytest.raises(ValueError):
        c.fft(np.zeros((3, 4), dtype=np.uint64))                             # not a power of two
    with pytest.raises(ValueError):
        c.msm(srs[:4], oracle.from_ints([1, 2, 3]))                          # length mismatch


def test_emu_lagrange_srs_and_evaluation_commitments(emu, oracle, golden):
    """Lagrange-basis SRS against the defining products of the big-int model, and the commitments taken from the wire VALUES
    against it equal the commitments of the wire POLYNOMIALS against the monomial powers of the same beta (and the oracle's)."""
    from oracle import pymodel as pm
    c0 = pg.StandardComposer(_cdll=emu)
    beta_i = 0x1d0c3a5e7f9b2468ace013579bdf02468ace13579bdf048c159d26ae37bf48c1 % Q
    beta = oracle.from_ints([beta_i])
    for log_n in (0, 1, 3, 4):
        got = oracle.g1_to_ints(to_oracle(c0.srs_lagrange(beta[0], log_n)))
        assert got == pm.srs_lagrange(beta_i, log_n), log_n
    with pytest.raises(pg.EngineError):
        c0.srs_lagrange(oracle.from_ints([pm.group_gen(3)])[0], 3)            # beta on the domain
    for name in ("kat_range_check_0_ok", "batch_is_non_zero_maybe_equal", "kat_max_bound_0_ok"):
        _s, oc = run_oracle(golden[name]["program"], return_composer=True)
        _snap, c = run_engine(golden[name]["program"], lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
        k = c.domain_log_size()
        mono, lag = c.srs_powers(beta[0], 1 << k), c.srs_lagrange(beta[0], k)
        from_values = c.commit_wire_evaluations(lag)
        assert np.array_equal(from_values, c.commit_wire_polynomials(mono)), name
        polys = oc.wire_polynomials()
        for w in range(4):
            assert np.array_equal(to_oracle(from_values[w:w + 1]), oracle.g1_msm(to_oracle(mono), polys[w])), (name, w)
        with pytest.raises(pg.EngineError):
            c.commit_wire_evaluations(lag, log_n=k + 1)                       # a Lagrange SRS belongs to one domain size


def test_emu_msm_skewed_scalars(emu, oracle):
    """Scalars as wire values have them: tens of thousands of ones (one bucket whose run spans many parts and more than one
    group), small values, zeros.  Expected value by linearity: per distinct point, the sum of its scalars (one small MSM in the
    oracle)."""
    c = pg.StandardComposer(_cdll=emu)
    base = oracle.srs_powers(oracle.from_ints([0xabcdef]), 48)
    n = 90000                                                    # ~67 000 ones: 131 parts of <= 512 entries, 3 groups of <= 64 parts
    idx = np.arange(n) % 48
    sc = [1] * n
    for i in range(0, n, 7):
        sc[i] = 0
    for i in range(3, n, 11):
        sc[i] = (i * 2654435761) % 5
    for i in range(5, n, 501):
        sc[i] = synth_wide(77, 1)[0]
    pts = to_engine(base)[idx]
    got = c.msm(pts, oracle.from_ints(sc))
    per_point = [sum(sc[i] for i in range(j, n, 48)) % Q for j in range(48)]
    assert np.array_equal(to_oracle(got.reshape(1, 12)), oracle.g1_msm(base, oracle.from_ints(per_point)))
