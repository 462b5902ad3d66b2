"""The sharded path on real GPUs, through the C ABI: pg_comm_init (NCCL), pg_check_sharded (verdict all-reduce with the
sequential composer's row numbering), pg_gather_column / pg_gather_variables (gather of result and witness shards).
World of one rank on any B200 box; world of two when the box has two GPUs (gpurun --gpus 2) -- the gathered table must equal
the single-GPU sequential composer bit for bit."""
import os
import sys

import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import sharding

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sequential(oracle, circuit, device=0):
    c = pg.StandardComposer(device=device)
    outs, n_err = sharding.run_circuit(c, circuit)
    bad, first = c.check_circuit_satisfied()
    return c, outs, (bad, first, n_err)


@pytest.mark.parametrize("mode,fused", [(pg.CHECK_GENERIC, False), (pg.CHECK_SPARSE, False), (pg.CHECK_SPARSE, True)])
def test_world_of_one_over_nccl(oracle, mode, fused):
    from tests.test_shard_plan import mixed_circuit
    circuit = mixed_circuit(oracle, scale=40, wrong_claims=(3, 200))
    ref, ref_outs, verdict = _sequential(oracle, circuit)
    assert verdict[0] == 3 and verdict[2] == 1
    c = pg.StandardComposer(device=0, check_mode=mode, fused_check=fused)
    c.comm_init(pg.comm_unique_id(), 0, 1)
    mine = sharding.plan_of(circuit, 1, pg.SHARD_ROWS)[0]
    outs, n_err = sharding.run_circuit(c, circuit, mine)
    assert c.check_sharded(mine, n_err) == verdict
    assert c.check_sharded(None, n_err) == verdict
    y = np.empty((circuit[1]["n"], 4), dtype=np.uint64)
    total, counts = c.gather_column(outs[1], y)
    assert total == circuit[1]["n"] and counts[0] == total and (y == ref_outs[1].values()).all()
    for call in (0, 1, 4, 9):
        s = mine[call]
        cnt = (s.inst_hi - s.inst_lo) * pg.op_shape(circuit[call]["gadget"], circuit[call].get("num_bits", 0))[1]
        buf = np.empty((cnt, 4), dtype=np.uint64)
        assert c.gather_variables(call, buf) == cnt
        assert (buf == ref.variables(s.var_base, cnt)).all()
    with pytest.raises(pg.EngineError):
        c.gather_variables(1, np.empty((3, 4), dtype=np.uint64))           # destination too small
    c.comm_destroy()
    c.close(); ref.close()


def _rank(rank, world, uid, policy, mode, fused, q):
    sys.path.insert(0, ROOT)
    import torch
    from oracle import binding as ob
    from tests.test_shard_plan import mixed_circuit
    circuit = mixed_circuit(ob, scale=40, wrong_claims=(3, 200))
    c = pg.StandardComposer(device=rank, check_mode=mode, fused_check=fused)
    c.comm_init(uid, rank, world)
    mine = sharding.plan_of(circuit, world, policy)[rank]
    outs, n_err = sharding.run_circuit(c, circuit, mine)
    verdict = c.check_sharded(mine, n_err)
    dev = torch.device("cuda", rank)
    y = torch.empty((circuit[1]["n"], 4), dtype=torch.int64, device=dev)     # device destination: NCCL writes it directly
    total, counts = c.gather_column(outs[1], y)
    c.sync()
    tables = {}
    for call in (0, 1, 4, 6, 9, 11):
        cnt = circuit[call]["n"] * pg.op_shape(circuit[call]["gadget"], circuit[call].get("num_bits", 0))[1]
        buf = np.empty((cnt, 4), dtype=np.uint64)
        assert c.gather_variables(call, buf) == cnt
        tables[call] = buf
    q.put((rank, verdict, total, counts[:world], y.cpu().numpy().view(np.uint64), tables))
    c.comm_destroy()
    c.close()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("policy,mode,fused", [(0, 0, False), (1, 1, False), (0, 1, True)])
def test_two_gpus_equal_the_sequential_composer(oracle, policy, mode, fused):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from tests.test_shard_plan import mixed_circuit
    circuit = mixed_circuit(oracle, scale=40, wrong_claims=(3, 200))
    ref, ref_outs, verdict = _sequential(oracle, circuit)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid = pg.comm_unique_id()
    procs = [ctx.Process(target=_rank, args=(r, 2, uid, policy, mode, fused, q)) for r in range(2)]
    for p in procs: p.start()
    got = [q.get(timeout=600) for _ in range(2)]
    for p in procs: p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    plan = sharding.plan_of(circuit, 2, policy)
    for rank, v, total, counts, y, tables in got:
        assert v == verdict                                                  # every rank holds the verdict of the whole circuit
        assert total == circuit[1]["n"] and counts == [plan[r][1].inst_hi - plan[r][1].inst_lo for r in range(2)]
        assert (y == ref_outs[1].values()).all()
        for call, buf in tables.items():
            assert (buf == ref.variables(plan[0][call].var_base, buf.shape[0])).all(), call
    ref.close()
