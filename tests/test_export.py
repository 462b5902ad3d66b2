"""pg_export_composer: the import adapter's input.  The export of a mixed composer is replayed into the oracle's
composer through public StandardComposer methods only (add_input / poly_gate / range_gate) and must give the composer the oracle
builds by running the gadgets themselves: same Variables, wires, selector columns, dense PI, permutation."""
import ctypes as C
import os

import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import _lib
from tests.export_replay import replay
from tests.programs import synth_wide


def _build_both(make_composer, ob):
    n = 37
    x = [v if i % 3 else v % 2 ** 64 for i, v in enumerate(synth_wide(81, n))]
    s = [v & 1 for v in synth_wide(82, n)]
    claims = [1 if i % 3 == 0 else 0 for i in range(n)]
    c = make_composer()
    w = c.add_input(ob.from_ints(x)); sv = c.add_input(ob.from_ints(s))
    y = pg.range_check(c, ob.from_ints([0]), ob.from_ints([2 ** 64]), w)
    c.constrain_to_constant(y, ob.from_ints(claims), pi=ob.from_ints([7 * i for i in range(n)]))
    mb, _ = pg.max_bound(c, ob.from_ints([2 ** 20 + 5]), sv)
    so = pg.conditionally_select_one(c, w, sv)
    pg.conditionally_select_zero(c, so, sv)
    pg.maybe_equal(c, so, w)
    c.range_gate(sv, 4)
    pg.is_non_zero_flags(c, so, so.values())
    oc = ob.Composer()
    ow = oc.add_input_batch(ob.from_ints(x)); osv = oc.add_input_batch(ob.from_ints(s))
    oy = oc.range_check_batch(ob.from_ints([0]), ob.from_ints([2 ** 64]), ow)
    oc.constrain_to_constant_batch(oy, ob.from_ints(claims), ob.from_ints([7 * i for i in range(n)]))
    oc.max_bound_batch(ob.from_ints([2 ** 20 + 5]), osv)
    oso = oc.select_one_batch(ow, osv)
    oc.select_zero_batch(oso, osv)
    oc.maybe_equal_batch(oso, ow)
    oc.range_gate_batch(osv, 4)
    assert oc.is_non_zero_batch(oso, oc.variables()[oso.astype(np.int64)]) == (0, n)
    return c, oc


def _check(c, oc, ob, tmp_path, chunk_rows):
    path = os.path.join(tmp_path, "composer.pgexp")
    c.export(path, chunk_rows=chunk_rows, sigma=True)
    ex, rc = replay(path, ob)
    assert (ex.n_rows, ex.n_vars) == (oc.n, oc.n_vars) == (rc.n, rc.n_vars)
    assert (rc.variables() == oc.variables()).all()
    assert (rc.wires() == oc.wires()).all() and (ex.w_idx == oc.wires()).all()
    assert (rc.selectors() == oc.selectors()).all()
    assert (rc.dense_pi() == oc.dense_pi()).all()
    assert rc.check() == oc.check()
    # the exported permutation against the oracle's variable_map: successor of every position inside its Variable's cycle
    for var in (0, 5, 5 + 37, int(ex.w_idx[2, 40]), oc.n_vars - 1):
        uses = oc.perm_of(var)
        for (row, wire), (nrow, nwire) in zip(uses, uses[1:] + uses[:1]):
            assert ex.sigma[wire, row] == nrow * 4 + nwire


def test_export_replays_into_the_oracle_composer_emu(oracle, tmp_path):
    from tests.test_emu_engine import _build
    emu = _lib.bind(C.CDLL(_build("libpg_emu.so", "engine_emu.cpp")))
    c, oc = _build_both(lambda: pg.StandardComposer(_cdll=emu), oracle)
    _check(c, oc, oracle, str(tmp_path), chunk_rows=1000)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk_rows", [0, 777])
def test_export_replays_into_the_oracle_composer_gpu(oracle, tmp_path, chunk_rows):
    c, oc = _build_both(lambda: pg.StandardComposer(device=0), oracle)
    _check(c, oc, oracle, str(tmp_path), chunk_rows)
