"""pg_template_get (pure host code of the shipped library, no GPU): the rows one instance of a gadget appends -- wire references and
selector values -- must be what the oracle's composer holds after running that gadget once on a fresh composer."""
import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import _lib

Z = 0


def _resolve(w_ref, first_var, operands):
    out = np.zeros(w_ref.shape, dtype=np.uint64)
    for w in range(4):
        for r in range(w_ref.shape[1]):
            v = int(w_ref[w, r])
            out[w, r] = Z if v == 0 else first_var + v - 1 if v > 0 else operands[-1 - v]
    return out


@pytest.mark.parametrize("case", ["range_check", "max_bound", "maybe_equal", "is_non_zero", "select_zero", "select_one", "constrain", "range_gate"])
def test_template_matches_the_oracle_composer(oracle, case):
    lib = _lib.load()
    f = oracle.from_ints
    oc = oracle.Composer()
    x = oc.add_input_batch(f([123456])); y = oc.add_input_batch(f([1]))
    row0, var0 = oc.n, oc.n_vars
    a = b = None; bits = 0
    if case == "range_check":
        a, b = f([50000])[0], f([250000])[0]
        oc.range_check_batch(f([50000]), f([250000]), x); gadget, ops = pg.OP_RANGE_CHECK, [int(x[0])]
    elif case == "max_bound":
        b = f([2 ** 100 + 7])[0]
        oc.max_bound_batch(f([2 ** 100 + 7]), x); gadget, ops = pg.OP_MAX_BOUND, [int(x[0])]
    elif case == "maybe_equal":
        oc.maybe_equal_batch(x, y); gadget, ops = pg.OP_MAYBE_EQUAL, [int(x[0]), int(y[0])]
    elif case == "is_non_zero":
        assert oc.is_non_zero_batch(x, f([123456])) == (0, 1); gadget, ops = pg.OP_IS_NON_ZERO, [int(x[0])]
    elif case == "select_zero":
        oc.select_zero_batch(x, y); gadget, ops = pg.OP_SELECT_ZERO, [int(x[0]), int(y[0])]
    elif case == "select_one":
        oc.select_one_batch(x, y); gadget, ops = pg.OP_SELECT_ONE, [int(x[0]), int(y[0])]
    elif case == "constrain":
        a, b = f([77])[0], f([123456])[0]                      # (pi, constant)
        oc.constrain_to_constant_batch(x, f([123456]), f([77])); gadget, ops = pg.OP_CONSTRAIN, [int(x[0])]
    else:
        bits = 20
        oc.range_gate_batch(x, bits); gadget, ops = pg.OP_RANGE_GATE, [int(x[0])]
    w_ref, sel, gate, n_vars = pg.template_get(gadget, bits, a, b, _cdll=lib)
    rows = w_ref.shape[1]
    assert (rows, n_vars) == (oc.n - row0, oc.n_vars - var0)
    k = 0
    if case in ("range_check", "max_bound"):
        k = (rows - 11) // 4 if case == "range_check" else (rows - 5) // 2
    assert (rows, n_vars) == pg.op_shape(gadget, k or bits)
    assert (_resolve(w_ref, var0, ops) == oc.wires()[:, row0:]).all()
    assert (sel == oc.selectors()[:6, row0:]).all()
    q_arith, q_range = oc.selectors()[6, row0:], oc.selectors()[7, row0:]
    one = f([1])[0]
    assert ((gate == 0) == (q_arith == one).all(axis=1)).all() and ((gate == 1) == (q_range == one).all(axis=1)).all()


def test_template_get_rejects_bad_arguments():
    lib = _lib.load()
    with pytest.raises(ValueError):
        pg.template_get(pg.OP_RANGE_CHECK, 0, None, None, _cdll=lib)            # bounds missing
    with pytest.raises(ValueError):
        pg.template_get(pg.OP_RANGE_GATE, 7, _cdll=lib)                         # odd width
    with pytest.raises(ValueError):
        pg.template_get(99, 0, _cdll=lib)
