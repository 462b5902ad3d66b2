"""CPU tests of the evaluation-domain path (pg_fft, pg_wire_polynomials) through tests/emu: the engine's pass planning,
tile indexing, twiddle addressing and bit-reversal (ntt.cuh) run on the loop backend and are compared with the oracle.
The CUDA kernel k_ntt_pass itself is checked on the B200 by tests/test_gpu_parity.py."""
import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from tests.engine_runner import run_engine
from tests.programs import Q, hx, run_oracle, synth_wide
from tests.test_emu_engine import emu  # noqa: F401  (fixture)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 9, 11, 12, 13])
def test_emu_fft_matches_oracle(emu, oracle, log_n):
    c = pg.StandardComposer(_cdll=emu)
    a = oracle.from_ints(synth_wide(70 + log_n, 1 << log_n))
    assert np.array_equal(c.fft(a), oracle.fft(a))
    assert np.array_equal(c.fft(a, inverse=True), oracle.fft(a, inverse=True))
    assert np.array_equal(c.fft(c.fft(a), inverse=True), a)


def test_emu_fft_golden(emu, oracle, golden_domain):
    c = pg.StandardComposer(_cdll=emu)
    for log_n, vec in golden_domain["fft"].items():
        a = oracle.from_ints([int(x, 16) for x in vec["input"]])
        assert [hx(v) for v in oracle.to_ints(c.fft(a))] == vec["fft"]
        assert [hx(v) for v in oracle.to_ints(c.fft(a, inverse=True))] == vec["ifft"]


def test_emu_pass_plan_covers_every_stage(emu, oracle):
    """Sizes whose stages split into 2 and 3 passes with uneven stage counts (2^14: 11+3, 2^20: 11+5+4 on the loop backend is
    too slow, so the 3-pass case is forced with a small maximum tile through the same planner in test_gpu_parity)."""
    c = pg.StandardComposer(_cdll=emu)
    a = oracle.from_ints(synth_wide(88, 1 << 14))
    assert np.array_equal(c.fft(a), oracle.fft(a))


def test_emu_wire_polynomials(emu, oracle, golden, golden_domain):
    from oracle.gen_golden import coeff_digest
    n = 0
    for name, spec in golden.items():
        if spec["expected"]["error"]:
            continue
        _s, oc = run_oracle(spec["program"], return_composer=True)
        _snap, c = run_engine(spec["program"], lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
        got = c.wire_polynomials()
        assert c.domain_log_size() == (oc.n - 1).bit_length()
        assert np.array_equal(got, oc.wire_polynomials()), name
        if name in golden_domain["wire_polynomials"]:
            assert coeff_digest([oracle.to_ints(got[w]) for w in range(4)]) == golden_domain["wire_polynomials"][name]["digest"]
            n += 1
    assert n >= 4


def test_emu_wire_polynomials_domain_argument(emu, oracle, golden):
    spec = golden["kat_range_check_0_ok"]
    _snap, c = run_engine(spec["program"], lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
    k = c.domain_log_size()
    # a larger domain is allowed (more zero padding): same polynomial values on the rows
    big = c.wire_polynomials(k + 2)
    assert big.shape == (4, 1 << (k + 2), 4)
    back = oracle.fft(big[1])                                  # evaluations over the larger domain
    rows = c.rows()
    assert np.array_equal(back[: c.circuit_size()], rows["w_val"][1]) and not back[c.circuit_size():].any()
    with pytest.raises(pg.EngineError):
        c.wire_polynomials(k - 1)                              # domain smaller than the circuit
    with pytest.raises(pg.EngineError):
        c.wire_polynomials(33)                                 # beyond the field's two-adicity
