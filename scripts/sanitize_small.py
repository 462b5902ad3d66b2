"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel class on tiny batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plonk_gadgets_b200 as pg
from oracle import binding as ob
from tests.programs import synth_wide

n = 300
_cdll = None
if "--emu" in sys.argv:      # host-backend dry run of this script's logic (no GPU): tests/emu
    import ctypes
    from plonk_gadgets_b200 import _lib
    _cdll = _lib.bind(ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "emu", "_build", "libpg_emu.so")))
vals = synth_wide(3, n)
wit = [v if i % 2 else v % 2 ** 64 for i, v in enumerate(vals)]
for mode in (pg.CHECK_GENERIC, pg.CHECK_SPARSE):
    c = pg.StandardComposer(device=0, check_mode=mode, _cdll=_cdll)
    w = c.add_input(ob.from_ints(wit))
    y = pg.range_check(c, ob.from_ints([0]), ob.from_ints([2 ** 64]), w)
    mb, k = pg.max_bound(c, ob.from_ints([2 ** 200 + 1 + i for i in range(n)]), w)
    eq = pg.maybe_equal(c, y, mb)
    s1 = pg.conditionally_select_one(c, w, y)
    s0 = pg.conditionally_select_zero(c, s1, eq)
    pg.is_non_zero(c, s1, s1.values())
    c.constrain_to_constant(eq, ob.from_ints([1]), ob.from_ints(list(range(n))))
    bad, first = c.check_circuit_satisfied()
    rows = c.rows()
    v = c.variables()
    sig = c.permutation()
    raw = c.to_bytes(v[:100]); back, nb, _ = c.from_bytes(raw)
    assert nb == 0 and (back == v[:100]).all()
    inv = c.fr_op(4, v[:1000])
    print("mode", mode, "rows", c.circuit_size(), "vars", c.num_variables(), "unsat", bad, "sigma", sig.shape, "ok")
    c.close()
