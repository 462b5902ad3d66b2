"""Runs a test program (tests/programs.py) on the batched engine through the C ABI and snapshots the composer."""
from __future__ import annotations

import numpy as np

import plonk_gadgets_b200 as pg
from tests.programs import SEL_NAMES, Snapshot, _vals, unsat_rows


def run_engine(program, make_composer, ob, return_composer: bool = False):
    """`make_composer()` -> plonk_gadgets_b200.StandardComposer; `ob` = oracle.binding (only used to convert between
    canonical integers and Montgomery limbs -- the engine's verdict and dumps are its own)."""
    c = make_composer()
    cols, error = {}, None
    for idx, op in enumerate(program):
        kind = op["op"]
        if kind == "add_input":
            cols[idx] = c.add_input(ob.from_ints(_vals(op, "values")))
        elif kind == "range_check":
            cols[idx] = pg.range_check(c, ob.from_ints(_vals(op, "min")), ob.from_ints(_vals(op, "max")), cols[op["witness"]])
        elif kind == "max_bound":
            cols[idx], _k = pg.max_bound(c, ob.from_ints(_vals(op, "max")), cols[op["witness"]])
        elif kind == "maybe_equal":
            cols[idx] = pg.maybe_equal(c, cols[op["a"]], cols[op["b"]])
        elif kind == "is_non_zero":
            try:
                pg.is_non_zero(c, cols[op["var"]], ob.from_ints(_vals(op, "assigned")))
            except pg.NonExistingInverse as e:
                error = (idx, "NonExistingInverse", e.first_err)
                break
        elif kind == "select_zero":
            cols[idx] = pg.conditionally_select_zero(c, cols[op["x"]], cols[op["select"]])
        elif kind == "select_one":
            cols[idx] = pg.conditionally_select_one(c, cols[op["y"]], cols[op["select"]])
        elif kind == "range_gate":
            c.range_gate(cols[op["witness"]], int(op["num_bits"]))
        elif kind == "constrain_to_constant":
            pi = ob.from_ints(_vals(op, "pi")) if op.get("pi") is not None else None
            c.constrain_to_constant(cols[op["a"]], ob.from_ints(_vals(op, "constant")), pi)
        else:
            raise ValueError(kind)
    snap = snapshot_of_engine(c, cols, error, ob)
    return (snap, c) if return_composer else snap


def snapshot_of_engine(c, cols, error, ob) -> Snapshot:
    n_rows, n_vars = c.circuit_size(), c.num_variables()
    variables = ob.to_ints(c.variables())
    rows = c.rows()
    one = ob.from_ints([1])[0]
    sel = [ob.to_ints(rows["sel"][k]) for k in range(6)]
    # q_arith / q_range from the engine (pg_materialize_gate_selectors); q_logic and the two group-addition selectors are 0 on
    # every row (include/pg_b200.h, pg_materialize_rows)
    q_arith, q_range = c.gate_selectors()
    sel += [ob.to_ints(q_arith), ob.to_ints(q_range)] + [[0] * n_rows] * 3
    assert len(sel) == len(SEL_NAMES) and int(one[0]) != 0
    snap = Snapshot(n_rows, n_vars, variables, rows["w_idx"], sel, ob.to_ints(rows["pi"]), [],
                    {k: [int(x) for x in v.ids()] for k, v in cols.items()}, error)
    bad, first = c.check_circuit_satisfied()              # the engine's own verdict (CUDA gate-check kernel)
    snap.unsat = unsat_rows(snap)                         # big-int evaluation of the engine's dump
    assert bad == len(snap.unsat), (bad, snap.unsat[:8])
    assert first == (snap.unsat[0] if snap.unsat else None), (first, snap.unsat[:4])
    # wire values reported by the materialiser must be the variables the wire indices point at
    wv = rows["w_val"]
    for w in range(4):
        assert ob.to_ints(wv[w]) == [variables[int(i)] for i in rows["w_idx"][w]], f"wire values of column {w}"
    # and the stand-alone row checker must agree on the materialised rows
    bad2, first2 = c.check_rows(rows["w_val"], rows["sel"], rows["pi"], q_arith, q_range)
    assert (bad2, first2) == (bad, first), (bad2, first2, bad, first)
    # without the gate selectors it evaluates the arithmetic widget alone (q_arith = 1): same verdict on the arithmetic rows
    bad3, first3 = c.check_rows(rows["w_val"], rows["sel"], rows["pi"])
    arith_bad = [r for r in snap.unsat if sel[6][r]]
    assert (bad3, first3) == (len(arith_bad), arith_bad[0] if arith_bad else None), (bad3, first3, bad, first)
    return snap
