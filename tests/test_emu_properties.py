"""Property tests (hypothesis) of the engine's host logic + kernel bodies through the test-only host backend, against the C
oracle: random bit widths, random/boundary witnesses, random gadget compositions (L3 of SURVEY.md 4.5), and variable-level
fault injection through the stand-alone row checker (L4)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import _lib
from tests import property_cases as pc
from tests.engine_runner import run_engine
from tests.programs import Q, hx, run_oracle
from tests.test_emu_engine import _build


@pytest.fixture(scope="module")
def emu():
    return _lib.bind(C.CDLL(_build("libpg_emu.so", "engine_emu.cpp")))


scalars = st.one_of(st.integers(0, Q - 1), st.integers(0, 2 ** 64), st.sampled_from([0, 1, Q - 1, Q - 2, 2 ** 255 % Q]))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(bits=st.integers(1, 253), seed=st.integers(0, 2 ** 32), extra=st.lists(scalars, min_size=1, max_size=4))
def test_range_gadgets_random_bounds(emu, oracle, bits, seed, extra):
    rng = np.random.default_rng(seed)
    top = 1 << (bits - 1)
    mx = (int(rng.integers(0, 2 ** 62)) % top | top) + 1 if bits > 1 else 2       # max-1 has exactly `bits` bits
    mn = int(rng.integers(0, 2 ** 62)) % mx
    wit = [mn, mx - 1, mx, (mn - 1) % Q] + extra
    prog = [dict(op="add_input", values=[hx(x) for x in wit]), dict(op="range_check", min=hx(mn), max=hx(mx), witness=0),
            dict(op="max_bound", max=hx(mx), witness=0)]
    so = run_oracle(prog)
    se = run_engine(prog, lambda: pg.StandardComposer(_cdll=emu), oracle)
    assert se.digest() == so.digest()
    assert se.unsat == so.unsat == []
    r = se.results(1)
    assert r[0] == 1 and r[1] == 1 and r[2] == 0


ops = st.sampled_from(pc.OPS)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 5), vals=st.lists(scalars, min_size=10, max_size=10), seq=st.lists(ops, min_size=1, max_size=6), seed=st.integers(0, 1000))
def test_random_gadget_compositions(emu, oracle, n, vals, seq, seed):
    """Later calls consume the columns earlier calls produced; the whole composer must equal the oracle's."""
    prog = pc.small_composition(n, vals, seq, seed)
    from tests.programs import expected_sigma
    so, oc = run_oracle(prog, return_composer=True)
    se, c = run_engine(prog, lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
    assert se.error == so.error
    assert se.digest() == so.digest()
    assert se.unsat == so.unsat
    assert (c.permutation() == expected_sigma(oc)).all()          # copy-constraint cycles


@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(row=st.integers(3, 3 + 4 * 9 + 11 - 1), col=st.integers(0, 2), delta=st.integers(1, Q - 1))
def test_fault_injection_flips_exactly_its_row(emu, oracle, row, col, delta):
    c = pg.StandardComposer(_cdll=emu)
    w = c.add_input(oracle.from_ints([77]))
    pg.range_check(c, oracle.from_ints([10]), oracle.from_ints([200]), w)      # k = 9: 47 rows
    rows = c.rows()
    assert c.check_rows(rows["w_val"], rows["sel"], rows["pi"]) == (0, None)
    sel_names = ("q_l", "q_r", "q_o")
    sel_idx = {0: 1, 1: 2, 2: 3}[col]
    bad = rows["w_val"].copy()
    old = oracle.to_ints(bad[col, row][None])[0]
    bad[col, row] = oracle.from_ints([(old + delta) % Q])[0]
    n_bad, first = c.check_rows(bad, rows["sel"], rows["pi"])
    # the row changes by selector * delta (+ q_m * delta * other wire for the multiplicative wires)
    sel = oracle.to_ints(rows["sel"][:, row])
    others = [oracle.to_ints(rows["w_val"][k, row][None])[0] for k in range(3)]
    change = sel[sel_idx] * delta
    if col < 2:
        change += sel[0] * delta * others[1 - col]
    assert (n_bad, first) == ((1, row) if change % Q else (0, None)), sel_names[col]


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(vals=st.lists(scalars, min_size=8, max_size=8), seq=st.lists(ops, min_size=1, max_size=5), seed=st.integers(0, 1000))
def test_sparse_program_random_compositions(emu, oracle, vals, seq, seed):
    """Batches of 48 instances go through the compiled structure-aware row program (SparseProgBody): for random gadget
    sequences -- satisfied or not -- its list of violated rows must be the oracle's, like the generic evaluation's."""
    prog = pc.batch_composition(vals, seq, seed)
    so = run_oracle(prog)
    for mode in (pg.CHECK_SPARSE, pg.CHECK_GENERIC):
        se = run_engine(prog, lambda: pg.StandardComposer(check_mode=mode, _cdll=emu), oracle)
        assert se.error == so.error and se.digest() == so.digest() and se.unsat == so.unsat, mode


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(vals=st.lists(scalars, min_size=8, max_size=8), seq=st.lists(ops, min_size=1, max_size=5), seed=st.integers(0, 1000),
       n=st.sampled_from([5, 33, 48]), pokes=st.lists(st.tuples(st.integers(0, 2 ** 30), scalars), min_size=1, max_size=4))
def test_fused_verdict_record_survives_pokes(emu, oracle, vals, seq, seed, n, pokes):
    """PG_F_FUSED_CHECK: whatever is overwritten after generation (any Variable of any segment, also one whose segment had recorded
    unsatisfied rows) and whatever is called afterwards, pg_check returns the big-int verdict of the composer's own row dump."""
    from tests.fault_cases import eval_rows

    def bigint_verdict(c):
        bad = eval_rows(oracle, c.rows(0, c.circuit_size(), want=("w_val", "sel", "pi")))
        return (len(bad), bad[0] if bad else None)

    prog = pc.batch_composition(vals, seq, seed, n=n)
    se, c = run_engine(prog, lambda: pg.StandardComposer(check_mode=pg.CHECK_SPARSE, fused_check=True, _cdll=emu), oracle, return_composer=True)
    assert c.check_circuit_satisfied() == bigint_verdict(c)
    nv = c.num_variables()
    for where, value in pokes:
        try:
            c.poke_variable(5 + where % (nv - 5), oracle.from_ints([value])[0])
        except pg.EngineError:
            continue                                   # a packed bit variable: cannot be overwritten
        assert c.check_circuit_satisfied() == bigint_verdict(c)
    if se.error is None:                               # calls made after a poke are recorded afresh
        a = c.add_input(oracle.from_ints(vals[:4] + [0] * (n - 4)))
        pg.maybe_equal(c, a, a)
        pg.is_non_zero_flags(c, a, oracle.from_ints([v if i % 2 else 0 for i, v in enumerate(vals[:4] + [1] * (n - 4))]), pg.NZ_UNIFORM)
        assert c.check_circuit_satisfied() == bigint_verdict(c)
        try:
            c.poke_variable(nv + 1, oracle.from_ints([pokes[0][1]])[0])
            assert c.check_circuit_satisfied() == bigint_verdict(c)
        except pg.EngineError:
            pass
    c.close()
