"""Property tests (hypothesis) on the CUDA library: the random gadget compositions of tests/test_emu_properties.py -- now with
dusk-plonk's range gate among the operations -- run through the kernels themselves (row-parallel check for the small batches, the
compiled row program and the per-instance kernels for the larger ones, the fused check), against the C oracle: Variables, wire
indices, selector rows, public inputs, the list of violated rows and the copy-constraint map."""
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import plonk_gadgets_b200 as pg
from tests import property_cases as pc
from tests.engine_runner import run_engine
from tests.programs import Q, expected_sigma, run_oracle

pytestmark = pytest.mark.gpu

scalars = st.one_of(st.integers(0, Q - 1), st.integers(0, 2 ** 64), st.sampled_from([0, 1, Q - 1, Q - 2, 2 ** 255 % Q]))
ops = st.sampled_from(pc.OPS_WITH_RANGE_GATE)
MODES = [dict(check_mode=pg.CHECK_GENERIC), dict(check_mode=pg.CHECK_SPARSE), dict(check_mode=pg.CHECK_SPARSE, fused_check=True)]


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 5), vals=st.lists(scalars, min_size=10, max_size=10), seq=st.lists(ops, min_size=1, max_size=6), seed=st.integers(0, 1000))
def test_random_gadget_compositions_gpu(oracle, n, vals, seq, seed):
    prog = pc.small_composition(n, vals, seq, seed)
    so, oc = run_oracle(prog, return_composer=True)
    for kw in MODES:
        se, c = run_engine(prog, lambda: pg.StandardComposer(device=0, **kw), oracle, return_composer=True)
        assert se.error == so.error, kw
        assert se.digest() == so.digest(), kw
        assert se.unsat == so.unsat, kw
        assert (c.permutation() == expected_sigma(oc)).all(), kw
        c.close()


@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(vals=st.lists(scalars, min_size=8, max_size=8), seq=st.lists(ops, min_size=1, max_size=5), seed=st.integers(0, 1000),
       n=st.sampled_from([48, 33, 257]))
def test_batch_compositions_gpu(oracle, vals, seq, seed, n):
    prog = pc.batch_composition(vals, seq, seed, n=n)
    so = run_oracle(prog)
    for kw in MODES:
        se, c = run_engine(prog, lambda: pg.StandardComposer(device=0, **kw), oracle, return_composer=True)
        assert se.error == so.error and se.digest() == so.digest() and se.unsat == so.unsat, kw
        c.close()
