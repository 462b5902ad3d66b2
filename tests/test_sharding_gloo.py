"""Multi-process path on CPU: world_size-2 gloo.  Each rank runs its contiguous shard of one batched call (through the
test-only host backend of tests/emu -- there is no GPU here), and the verdict all-reduce / result all-gather of
plonk_gadgets_b200.sharding must reproduce the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import ctypes as C
    import plonk_gadgets_b200 as pg
    from plonk_gadgets_b200 import _lib, sharding
    from oracle import binding as ob
    from tests.programs import synth_wide
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = _lib.bind(C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libpg_emu.so")))
    vals = synth_wide(9, n)
    wit = [v if i % 2 else v % 2 ** 64 for i, v in enumerate(vals)]
    claims = [1 if i % 2 == 0 else 0 for i in range(n)]
    claims[5] ^= 1; claims[n - 3] ^= 1                       # two wrong claims -> two unsatisfied rows, in different shards
    lo, hi = sharding.shard_range(n, rank, world)
    c = pg.StandardComposer(_cdll=emu)
    w = c.add_input(ob.from_ints(wit[lo:hi]))
    y = pg.range_check(c, ob.from_ints([0]), ob.from_ints([2 ** 64]), w)
    bad, first = c.check_circuit_satisfied()
    assert (bad, first) == (0, None)
    # claims are checked as a second pass so that the gadget rows keep the uniform 271-row stride
    res = ob.to_ints(y.values())
    n_bad = sum(int(r != k) for r, k in zip(res, claims[lo:hi]))
    first_local = next((i for i, (r, k) in enumerate(zip(res, claims[lo:hi])) if r != k), None)
    off = sharding.shard_offsets(n, rank, world, 271, 653)
    first_global = None if first_local is None else sharding.FRESH_ROWS + off.row_offset + 271 * first_local + 270
    tot_bad, tot_first, tot_err = sharding.allreduce_verdict(n_bad, first_global, 0)
    gathered = sharding.allgather_ragged(torch.from_numpy(y.values().view(np.int64)))
    gathered = torch.cat(gathered, dim=0)
    # gather of witness shards: the two calls appended 1 and 653 variables per instance
    table = sharding.allgather_witness_shards(torch.from_numpy(c.variables().view(np.int64)), [hi - lo, 653 * (hi - lo)])
    # second call shape: the native range gate (10 rows per 64-bit witness), verdict straight from the engine's own check;
    # odd instances are uniform Fr and do not fit 64 bits, so both shards hold unsatisfied rows
    c2 = pg.StandardComposer(_cdll=emu)
    c2.range_gate(c2.add_input(ob.from_ints(wit[lo:hi])), 64)
    bad2, first2 = c2.check_circuit_satisfied()
    off2 = sharding.shard_offsets(n, rank, world, 10, 32)
    rg = sharding.allreduce_verdict(bad2, None if first2 is None else first2 + off2.row_offset, 0)
    if rank == 0:
        q.put((tot_bad, tot_first, tot_err, gathered.numpy().view(np.uint64).copy(), off.row_offset, off.var_offset, rg,
               table.numpy().view(np.uint64).copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_single_run(oracle):
    import ctypes as C
    import plonk_gadgets_b200 as pg
    from plonk_gadgets_b200 import _lib
    from tests.programs import synth_wide
    from tests.test_emu_engine import _build
    emu = _lib.bind(C.CDLL(_build("libpg_emu.so", "engine_emu.cpp")))
    n = 17                                                   # ragged: 8 + 9 instances
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs: p.start()
    tot_bad, tot_first, tot_err, gathered, off_r, off_v, rg, table = q.get(timeout=240)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    # single-process run of the whole batch
    vals = synth_wide(9, n)
    wit = [v if i % 2 else v % 2 ** 64 for i, v in enumerate(vals)]
    c = pg.StandardComposer(_cdll=emu)
    w = c.add_input(oracle.from_ints(wit))
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    assert (gathered == y.values()).all()
    assert table.shape == (5 + n * 654, 4) and (table == c.variables()).all()    # the sequential composer's variable table
    assert (tot_bad, tot_err) == (2, 0)
    assert tot_first == 3 + 271 * 5 + 270                    # the output row of instance 5 in the sequential numbering
    assert (off_r, off_v) == (0, 0)
    c2 = pg.StandardComposer(_cdll=emu)
    c2.range_gate(c2.add_input(oracle.from_ints(wit)), 64)
    assert rg == c2.check_circuit_satisfied() + (0,) == (n // 2, 3 + 10 + 9, 0)


def test_shard_ranges_cover_everything():
    from plonk_gadgets_b200 import sharding
    for n in (0, 1, 7, 8, 1000, 2 ** 24):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1
