"""Sharding plan (pg_shard_plan, pure host code of the shipped library) and the sharded path on CPU: world_size-2 gloo, each
rank replaying its part of a mixed circuit (BASELINE config C5 in miniature) through the test-only host backend, verdict and
witness shards combined -- the union must be the SEQUENTIAL composer of the whole circuit, bit for bit.
On GPUs the two collectives are NCCL behind the C ABI (pg_check_sharded / pg_gather_*: tests/test_gpu_multi.py)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import _lib, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mixed_circuit(ob, scale=1, wrong_claims=()):
    """range_check k=65 / max_bound k=253 / is_non_zero / select_one + select_zero / maybe_equal, with a claim row per range_check
    instance; `wrong_claims`: range_check instances whose claim is flipped (=> one unsatisfied row each)."""
    from tests.programs import synth_wide
    n_rc, n_mb, n_nz, n_sel = 11 * scale, 5 * scale, 13 * scale, 9 * scale
    x_rc = [v if i % 2 else v % 2 ** 64 for i, v in enumerate(synth_wide(71, n_rc))]
    claims = [(1 if i % 2 == 0 else 0) ^ (1 if i in wrong_claims else 0) for i in range(n_rc)]
    x_mb = [v % 2 ** 250 for v in synth_wide(72, n_mb)]
    x_nz = synth_wide(73, n_nz)
    x_nz[4] = 0                                               # one NonExistingInverse
    x_sel, s_sel = synth_wide(74, n_sel), [v & 1 for v in synth_wide(75, n_sel)]
    f = ob.from_ints
    return [
        {"gadget": pg.OP_ADD_INPUT, "n": n_rc, "group": 0, "values": f(x_rc)},
        {"gadget": pg.OP_RANGE_CHECK, "num_bits": 65, "n": n_rc, "group": 0, "min": f([0]), "max": f([2 ** 64]), "witness": 0},
        {"gadget": pg.OP_CONSTRAIN, "n": n_rc, "group": 0, "a": 1, "constant": f(claims)},
        {"gadget": pg.OP_ADD_INPUT, "n": n_mb, "group": 1, "values": f(x_mb)},
        {"gadget": pg.OP_MAX_BOUND, "num_bits": 253, "n": n_mb, "group": 1, "max": f([2 ** 252]), "witness": 3},
        {"gadget": pg.OP_ADD_INPUT, "n": n_nz, "group": 2, "values": f(x_nz)},
        {"gadget": pg.OP_IS_NON_ZERO, "n": n_nz, "group": 2, "var": 5, "assigned": f(x_nz)},
        {"gadget": pg.OP_ADD_INPUT, "n": n_sel, "group": 3, "values": f(x_sel)},
        {"gadget": pg.OP_ADD_INPUT, "n": n_sel, "group": 3, "values": f(s_sel)},
        {"gadget": pg.OP_SELECT_ONE, "n": n_sel, "group": 3, "x": 7, "select": 8},
        {"gadget": pg.OP_SELECT_ZERO, "n": n_sel, "group": 3, "x": 9, "select": 8},
        {"gadget": pg.OP_MAYBE_EQUAL, "n": n_sel, "group": 3, "a": 9, "b": 10},
    ]


def test_plan_is_a_partition_with_prefix_sum_bases():
    lib = _lib.load()
    ops = [(pg.OP_ADD_INPUT, 0, 1000, 0), (pg.OP_RANGE_CHECK, 65, 1000, 0), (pg.OP_ADD_INPUT, 0, 301, 1), (pg.OP_MAX_BOUND, 253, 301, 1),
           (pg.OP_ADD_INPUT, 0, 77, 2), (pg.OP_IS_NON_ZERO, 0, 77, 2), (pg.OP_RANGE_GATE, 64, 77, 2), (pg.OP_ADD_INPUT, 0, 5, 3)]
    shapes = [pg.op_shape(g, k) for g, k, _, _ in ops]
    assert shapes[1] == (271, 653) and shapes[3] == (511, 514) and shapes[6] == (10, 32)
    for policy in (pg.SHARD_EVEN, pg.SHARD_ROWS):
        for world in (1, 2, 3, 4, 8):
            plan = pg.shard_plan(ops, world, policy, _cdll=lib)
            row, var = 3, 5
            for k, (g, bits, n, grp) in enumerate(ops):
                r, v = shapes[k]
                assert plan[0][k].inst_lo == 0 and plan[-1][k].inst_hi == n
                for rank in range(world):
                    s = plan[rank][k]
                    assert s.inst_lo <= s.inst_hi
                    if rank + 1 < world:
                        assert s.inst_hi == plan[rank + 1][k].inst_lo
                    assert (s.row_base, s.var_base) == (row + s.inst_lo * r, var + s.inst_lo * v)      # the sequential composer's numbering
                    same = [j for j in range(len(ops)) if ops[j][3] == grp]
                    assert all((plan[rank][j].inst_lo, plan[rank][j].inst_hi) == (s.inst_lo, s.inst_hi) for j in same)
                row += n * r; var += n * v
            rows_of = [sum((plan[rank][k].inst_hi - plan[rank][k].inst_lo) * shapes[k][0] for k in range(len(ops))) for rank in range(world)]
            assert sum(rows_of) == row - 3
            if policy == pg.SHARD_ROWS:                                       # balanced to within one instance of the largest gadget
                assert max(rows_of) - min(rows_of) <= 2 * 511
    with pytest.raises(ValueError):
        pg.shard_plan([(pg.OP_ADD_INPUT, 0, 10, 0), (pg.OP_RANGE_CHECK, 65, 11, 0)], 2, _cdll=lib)    # one group, two lengths
    with pytest.raises(ValueError):
        pg.shard_plan([(pg.OP_RANGE_CHECK, 1, 10, 0)], 2, _cdll=lib)                                  # num_bits out of range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, policy, q):
    sys.path.insert(0, ROOT)
    from oracle import binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = _lib.bind(C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libpg_emu.so")))
    circuit = mixed_circuit(ob, wrong_claims=(2, 9))
    plan = sharding.plan_of(circuit, world, policy, _cdll=emu)
    mine = plan[rank]
    c = pg.StandardComposer(_cdll=emu)
    _, n_err = sharding.run_circuit(c, circuit, mine)
    # the engine numbers this rank's rows as the sequential composer does (pg_check_sharded; the host backend is a world of one,
    # so the cross-rank reduction is the gloo one here and NCCL on the GPUs)
    bad, first, err = c.check_sharded(mine, n_err)
    tot = sharding.allreduce_verdict(bad, first, err)
    # gather of witness shards, call by call, into the sequential composer's Variable order
    local = torch.from_numpy(c.variables().view(np.int64))
    sizes = [(s.inst_hi - s.inst_lo) * pg.op_shape(cc["gadget"], cc.get("num_bits", 0))[1] for s, cc in zip(mine, circuit)]
    table = sharding.allgather_witness_shards(local, sizes)
    if rank == 0:
        q.put((tot, table.numpy().view(np.uint64).copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("policy", [0, 1])
def test_two_rank_mixed_circuit_equals_sequential_composer(oracle, policy):
    from tests.test_emu_engine import _build
    emu = _lib.bind(C.CDLL(_build("libpg_emu.so", "engine_emu.cpp")))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, policy, q)) for r in range(2)]
    for p in procs: p.start()
    tot, table = q.get(timeout=500)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    circuit = mixed_circuit(oracle, wrong_claims=(2, 9))
    c = pg.StandardComposer(_cdll=emu)
    _, n_err = sharding.run_circuit(c, circuit)
    bad, first = c.check_circuit_satisfied()
    assert n_err == 1 and bad == 3                             # two wrong claims + the errored is_non_zero instance's last row
    assert tot == (bad, first, n_err)
    assert (table == c.variables()).all()
    # and against the oracle composer of the whole circuit
    oc = oracle.Composer()
    w = oc.add_input_batch(circuit[0]["values"])
    y = oc.range_check_batch(circuit[1]["min"], circuit[1]["max"], w)
    assert (table[: oc.n_vars] == oc.variables()).all()


# ---- properties of the plan over random circuits (hypothesis; the same examples on every run unless HYPOTHESIS_PROFILE=explore) ----
from hypothesis import given, settings, strategies as st  # noqa: E402

_GADGETS = [(pg.OP_ADD_INPUT, 0), (pg.OP_RANGE_CHECK, 65), (pg.OP_RANGE_CHECK, 2), (pg.OP_MAX_BOUND, 253), (pg.OP_MAYBE_EQUAL, 0),
            (pg.OP_IS_NON_ZERO, 0), (pg.OP_SELECT_ZERO, 0), (pg.OP_SELECT_ONE, 0), (pg.OP_CONSTRAIN, 0), (pg.OP_RANGE_GATE, 64)]


@settings(max_examples=60, deadline=None)
@given(groups=st.lists(st.tuples(st.integers(0, 5000), st.lists(st.sampled_from(_GADGETS), min_size=1, max_size=4)), min_size=1, max_size=6),
       world=st.integers(1, 9), policy=st.sampled_from([0, 1]), interleave=st.booleans())
def test_plan_properties_random_circuits(groups, world, policy, interleave):
    ops = []
    for g, (n, gadgets) in enumerate(groups):
        ops += [(gad, bits, n, g) for gad, bits in gadgets]
    if interleave:                                           # a group's calls need not be adjacent
        ops = ops[::2] + ops[1::2]
    plan = pg.shard_plan(ops, world, policy, _cdll=_lib.load())
    shapes = [pg.op_shape(gad, bits) for gad, bits, _, _ in ops]
    row, var = 3, 5
    for k, (gad, bits, n, grp) in enumerate(ops):
        r, v = shapes[k]
        assert plan[0][k].inst_lo == 0 and plan[-1][k].inst_hi == n
        for rank in range(world):
            s = plan[rank][k]
            assert s.inst_lo <= s.inst_hi <= n
            assert rank + 1 == world or s.inst_hi == plan[rank + 1][k].inst_lo               # ranges tile [0, n) in rank order
            assert (s.row_base, s.var_base) == (row + s.inst_lo * r, var + s.inst_lo * v)      # prefix sums = the sequential composer
            first = next(j for j in range(len(ops)) if ops[j][3] == grp)
            assert (s.inst_lo, s.inst_hi) == (plan[rank][first].inst_lo, plan[rank][first].inst_hi)   # one cut per group
        row += n * r; var += n * v
    if policy == pg.SHARD_ROWS:
        rows_of = [sum((plan[rank][k].inst_hi - plan[rank][k].inst_lo) * shapes[k][0] for k in range(len(ops))) for rank in range(world)]
        weight = {}
        for k, (_, _, _, grp) in enumerate(ops):
            weight[grp] = weight.get(grp, 0) + shapes[k][0]
        assert max(rows_of) - min(rows_of) <= 2 * max(weight.values())                        # balanced to within an instance of the heaviest group
