"""Evaluation-domain oracle: oracle/fft.c (restated serial radix-2 FFT) against the defining DFT sums of
oracle/pymodel.py and the committed fixtures tests/golden/domain.json."""
import random

from tests.programs import Q, hx, run_oracle, synth_wide


def test_root_of_unity_constant(oracle):
    """The recalled ROOT_OF_UNITY limbs are 7^((q-1)/2^32) in Montgomery form, a primitive 2^32-th root of unity."""
    from oracle import pymodel as pm
    assert (Q - 1) % (1 << 32) == 0 and ((Q - 1) >> 32) % 2 == 1
    w = pow(7, (Q - 1) >> 32, Q)
    assert w == pm.ROOT_OF_UNITY == oracle.domain_group_gen(32)
    limbs = [0xb9b58d8c5f0e466a, 0x5b1b4c801819d7ec, 0x0af53ae352a31e64, 0x5bf3adda19e9b27b]
    assert sum(l << (64 * i) for i, l in enumerate(limbs)) == w * pow(2, 256, Q) % Q
    assert pow(w, 1 << 32, Q) == 1 and pow(w, 1 << 31, Q) == Q - 1


def test_group_generators(oracle, golden_domain):
    from oracle import pymodel as pm
    for k, v in golden_domain["group_gen"].items():
        assert hx(oracle.domain_group_gen(int(k))) == v == hx(pm.group_gen(int(k)))
    assert oracle.domain_group_gen(0) == 1 and oracle.domain_group_gen(1) == Q - 1


def test_fft_golden(oracle, golden_domain):
    from oracle import pymodel as pm
    assert len(golden_domain["fft"]) >= 5
    for log_n, vec in golden_domain["fft"].items():
        vals = [int(x, 16) for x in vec["input"]]
        assert len(vals) == 1 << int(log_n) and vals == synth_wide(90 + int(log_n), len(vals))
        assert [hx(v) for v in pm.dft(vals)] == vec["fft"]                       # the fixture is current
        a = oracle.from_ints(vals)
        assert [hx(v) for v in oracle.to_ints(oracle.fft(a))] == vec["fft"]
        assert [hx(v) for v in oracle.to_ints(oracle.fft(a, inverse=True))] == vec["ifft"]


def test_fft_against_defining_sums(oracle):
    from oracle import pymodel as pm
    rnd = random.Random(5)
    for log_n in (0, 1, 4, 5, 7, 8):
        vals = [rnd.randrange(Q) for _ in range(1 << log_n)]
        vals[0] = 0; vals[-1] = Q - 1
        a = oracle.from_ints(vals)
        assert oracle.to_ints(oracle.fft(a)) == pm.dft(vals)
        assert oracle.to_ints(oracle.fft(a, inverse=True)) == pm.dft(vals, inverse=True)


def test_fft_properties(oracle):
    """Round trip, linearity, the transform of a delta and of a constant at 2^12."""
    n = 1 << 12
    x, y = synth_wide(95, n), synth_wide(96, n)
    fx, fy = oracle.to_ints(oracle.fft(oracle.from_ints(x))), oracle.to_ints(oracle.fft(oracle.from_ints(y)))
    assert oracle.to_ints(oracle.fft(oracle.from_ints(fx), inverse=True)) == x
    s = oracle.to_ints(oracle.fft(oracle.from_ints([(a + 3 * b) % Q for a, b in zip(x, y)])))
    assert s == [(a + 3 * b) % Q for a, b in zip(fx, fy)]
    assert oracle.to_ints(oracle.fft(oracle.from_ints([5] + [0] * (n - 1)))) == [5] * n
    assert oracle.to_ints(oracle.fft(oracle.from_ints([7] * n))) == [7 * n % Q] + [0] * (n - 1)


def test_wire_polynomials_golden(oracle, golden, golden_domain):
    from oracle import pymodel as pm
    from oracle.gen_golden import coeff_digest
    assert len(golden_domain["wire_polynomials"]) >= 4
    for name, exp in golden_domain["wire_polynomials"].items():
        _snap, c = run_oracle(golden[name]["program"], return_composer=True)
        polys = c.wire_polynomials()
        assert polys.shape == (4, 1 << exp["log_n"], 4)
        cols = [oracle.to_ints(polys[w]) for w in range(4)]
        assert coeff_digest(cols) == exp["digest"], name
        assert [[hx(v) for v in col[:3]] for col in cols] == exp["head"], name
        # evaluating the polynomial at w^r gives back the wire value of row r (and 0 on the padding rows)
        w = pm.group_gen(exp["log_n"])
        for r in (0, 1, c.n - 1, c.n):
            if r >= 1 << exp["log_n"]:
                continue
            x = pow(w, r, Q)
            val = sum(cf * pow(x, e, Q) for e, cf in enumerate(cols[0])) % Q
            want = _snap.variables[int(_snap.wires[0, r])] if r < c.n else 0
            assert val == want, (name, r)
