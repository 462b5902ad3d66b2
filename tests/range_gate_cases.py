"""Shared cases for dusk-plonk's native range gate (the path /root/reference/src/range.rs:9-12 recommends for
power-of-two bounds): run on the host-emulation backend by the CPU tests and on the B200 by the `-m gpu` tests."""
import random

import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from tests.engine_runner import run_engine
from tests.programs import Q, SEL_NAMES, Snapshot, hx, run_oracle, synth_wide, unsat_rows

WIDTHS = [2, 4, 6, 8, 10, 12, 14, 30, 32, 34, 62, 64, 66, 128, 250, 252, 254, 256]


def vs_oracle(make_composer, oracle, bits, modes=(pg.CHECK_GENERIC, pg.CHECK_SPARSE)):
    """Witnesses around 2^bits and uniform ones: full composer state (variables, wires, all selector columns) equal to the oracle's,
    unsatisfied rows = exactly the closing assert_equal rows of the witnesses that do not fit."""
    r = synth_wide(300 + bits, 8)
    top = min(2 ** bits, Q)
    wit = [0, 1, 3, top - 1, top % Q, (top + 1) % Q, r[0] % top, r[1] % top, r[2], r[3], Q - 1, (2 * top) % Q, r[4] % 2 ** (bits // 2 + 1)]
    prog = [dict(op="add_input", values=[hx(x) for x in wit]), dict(op="range_gate", witness=0, num_bits=bits)]
    so = run_oracle(prog)
    gates = (bits + 7) // 8
    assert so.n_rows == 3 + len(wit) * (gates + 2) and so.n_vars == 5 + len(wit) * (1 + bits // 2)
    assert so.unsat == [3 + i * (gates + 2) + gates + 1 for i, w in enumerate(wit) if w >= 2 ** bits]
    for mode in modes:
        se = run_engine(prog, lambda: make_composer(check_mode=mode), oracle)
        assert se.digest() == so.digest() and se.unsat == so.unsat, (bits, mode)


def bad_arguments(make_composer, oracle):
    c = make_composer()
    w = c.add_input(oracle.from_ints([7, 8]))
    for bits in (0, 1, 3, 63, 257, 258):
        with pytest.raises(pg.EngineError) as e:
            c.range_gate(w, bits)
        assert e.value.code == -2, bits                         # PG_ERR_ARG (the reference asserts on odd widths)
    with pytest.raises(pg.EngineError):
        c.range_gate(pg.Variables(c, 999, 2), 8)                # unknown column
    assert c.circuit_size() == 3 and c.num_variables() == 7    # nothing was appended
    e = c.add_input(np.empty((0, 4), dtype=np.uint64))
    c.range_gate(e, 16)                                          # empty batch: nothing appended either
    assert c.circuit_size() == 3 and c.check_circuit_satisfied() == (0, None)


def _snapshot_from_arrays(oracle, n_rows, variables, w_idx, sel6, qa, qr, pi):
    sel = [oracle.to_ints(sel6[k]) for k in range(6)] + [oracle.to_ints(qa), oracle.to_ints(qr)] + [[0] * n_rows] * 3
    assert len(sel) == len(SEL_NAMES)
    return Snapshot(n_rows, len(variables), variables, w_idx, sel, oracle.to_ints(pi), [])


def fault_injection(make_composer, oracle, seed=7, trials=24):
    """Change one materialised wire value of a range-gate circuit: pg_check_rows_ex must report exactly the rows a big-int
    evaluation of the same (arithmetic + range widget) equation reports -- a changed accumulator breaks its own gate and, when it
    sits on the fourth wire, the previous gate's D(d_next - 4a) term."""
    c = make_composer()
    vals = [v % 2 ** 40 for v in synth_wide(310, 5)]
    w = c.add_input(oracle.from_ints(vals))
    c.range_gate(w, 40)
    y = pg.maybe_equal(c, w, w)
    c.range_gate(y, 2)
    assert c.check_circuit_satisfied() == (0, None)
    rows = c.rows(); qa, qr = c.gate_selectors(); n = c.circuit_size()
    assert c.check_rows(rows["w_val"], rows["sel"], rows["pi"], qa, qr) == (0, None)
    range_rows = [r for r in range(n) if oracle.to_ints(qr[r:r + 1])[0] == 1]
    assert len(range_rows) == 5 * 5 + 5 * 1
    rng = random.Random(seed)
    seen_range_failure = False
    for t in range(trials):
        r_ = rng.choice(range_rows) if t % 2 == 0 else rng.randrange(3, n)
        col = rng.randrange(4)
        bad = rows["w_val"].copy()
        delta = rng.choice([1, 2, 3, 4, rng.randrange(1, Q)])
        old = oracle.to_ints(bad[col, r_:r_ + 1])[0]
        bad[col, r_] = oracle.from_ints([(old + delta) % Q])[0]
        # expected verdict: the equation over the changed dump, wire values taken as their own "variables"
        flat = [oracle.to_ints(bad[k]) for k in range(4)]
        variables = flat[0] + flat[1] + flat[2] + flat[3]
        w_idx = np.array([[k * n + i for i in range(n)] for k in range(4)], dtype=np.uint64)
        exp = unsat_rows(_snapshot_from_arrays(oracle, n, variables, w_idx, rows["sel"], qa, qr, rows["pi"]))
        got = c.check_rows(bad, rows["sel"], rows["pi"], qa, qr)
        assert got == (len(exp), exp[0] if exp else None), (t, r_, col, got, exp)
        seen_range_failure |= any(r in range_rows for r in exp)
    assert seen_range_failure


def poked_witness(make_composer, oracle, modes=(pg.CHECK_GENERIC, pg.CHECK_SPARSE), n=40, trials=10, seed=11, arith="max_bound"):
    """Overwrite stored variables (accumulators, the witness, a gadget result) of a circuit with range gates and arithmetic
    rows: pg_check (the kernels themselves) must report exactly the rows a big-int evaluation reports.  The circuit is satisfied
    before the change, so only rows that read the changed variable -- and, for a fourth wire, the row above (its d_next) -- are
    re-evaluated, from the engine's own dump of wires and selectors."""
    vals = [v % 2 ** 30 for v in synth_wide(320, n)]
    rng = random.Random(seed)

    def delta(f):
        return f * (f - 1) * (f - 2) * (f - 3)
    for mode in modes:
        c = make_composer(check_mode=mode)
        w = c.add_input(oracle.from_ints(vals))
        c.range_gate(w, 30)
        y = pg.max_bound(c, oracle.from_ints([2 ** 30]), w)[0] if arith == "max_bound" else pg.maybe_equal(c, w, w)
        c.range_gate(y, 2)
        assert c.check_circuit_satisfied() == (0, None)
        n_vars, n_rows = c.num_variables(), c.circuit_size()
        rows = c.rows(want=("w_idx", "sel", "pi")); qa, qr = c.gate_selectors()
        w_idx = rows["w_idx"]
        acc0 = 5 + n                                             # first accumulator of the first range gate (15 per instance)
        pokes = [acc0, acc0 + 14, acc0 + 15 * (n - 1) + 7, 5 + 3, int(y.ids()[2])]
        pokes += [rng.randrange(acc0, acc0 + 15 * n) for _ in range(trials)]
        pokes += [int(n_vars - 1 - rng.randrange(n)) for _ in range(2)]      # accumulators of the second range gate (one per instance)
        hit_range_row = False
        for var in pokes:
            old = c.variables(var, 1)[0].copy()
            new_val = (rng.choice([1, 2, 5, rng.randrange(Q)]) + oracle.to_ints(old[None])[0]) % Q
            c.poke_variable(var, oracle.from_ints([new_val])[0])
            cand = set()
            for k in range(4):
                for r in np.nonzero(w_idx[k] == var)[0]:
                    cand.add(int(r))
                    if k == 3 and r > 0:
                        cand.add(int(r) - 1)
            exp = []
            for r in sorted(cand):
                ids = [int(w_idx[k, r]) for k in range(4)] + [int(w_idx[3, (r + 1) % n_rows])]
                a, b, cc, d, dn = (new_val if v == var else oracle.to_ints(c.variables(v, 1))[0] for v in ids)
                q_m, q_l, q_r, q_o, q_4, q_c = (oracle.to_ints(rows["sel"][k, r:r + 1])[0] for k in range(6))
                pi, s_a, s_r = (oracle.to_ints(x[r:r + 1])[0] for x in (rows["pi"], qa, qr))
                g = s_a * (q_m * a * b + q_l * a + q_r * b + q_o * cc + q_4 * d + pi + q_c) \
                    + s_r * (delta(cc - 4 * d) + delta(b - 4 * cc) + delta(a - 4 * b) + delta(dn - 4 * a))
                if g % Q:
                    exp.append(r)
                    hit_range_row |= bool(s_r)
            assert exp, var
            assert c.check_circuit_satisfied() == (len(exp), exp[0]), (mode, var, exp)
            c.poke_variable(var, old)
            assert c.check_circuit_satisfied() == (0, None)
        assert hit_range_row
