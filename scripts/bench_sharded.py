"""BASELINE config C5 -- the mixed-gadget circuit of ~2^26 rows -- sharded over the GPUs of one box through the C ABI
(pg_shard_plan -> per-rank *_batch calls -> pg_check_sharded: the verdict all-reduce runs inside every timed step), plus the
bandwidth of the gather of witness shards (pg_gather_variables, NCCL over NVLink).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_sharded.py

Strong scaling: the circuit is fixed, each rank owns ~1/N of its rows.  One JSON line per (policy, check mode) on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import sharding

SEED = 0x706C6F6E6B5F6732
R2 = np.array([[0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]], dtype=np.uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2rows", type=int, default=26)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--gather-log2n", type=int, default=14, help="range_check instances per rank in the gather measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    def fresh_uid():
        """rank 0's ncclUniqueId, through torch.distributed (out of band for the C ABI); one id per communicator"""
        uid = torch.zeros(pg.api.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(pg.comm_unique_id()), dtype=torch.uint8).to(dev)
        if world > 1:
            dist.broadcast(uid, 0)
        return bytes(uid.cpu().numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for mode_name, mode in (("generic", pg.CHECK_GENERIC), ("sparse", pg.CHECK_SPARSE), ("fused", pg.CHECK_SPARSE)):
        c = pg.StandardComposer(device=local, check_mode=mode, timing=True, stream=stream.cuda_stream, fused_check=mode_name == "fused")
        c.comm_init(fresh_uid(), rank, world)
        to_mont = lambda ints: c.fr_op(0, np.array([[(v >> (64 * k)) & (2 ** 64 - 1) for k in range(4)] for v in ints], dtype=np.uint64), np.repeat(R2, len(ints), axis=0))
        b = to_mont([0, 2 ** 64, 2 ** 252])
        mn, mx, mx252 = b[0:1].copy(), b[1:2].copy(), b[2:3].copy()
        q = 1 << (args.log2rows - 2)
        n_rc, n_mb, n_nz, n_sel = q // 271, q // 511, q // 3, q // 5
        def synth(n, stream_id, kind, bits):
            t = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, stream_id, kind, bits, t); return t
        x_rc, x_mb, x_nz = synth(n_rc, 51, 2, 64), synth(n_mb, 52, 2, 250), synth(n_nz, 53, 0, 0)
        x_nz[1023::1024] = 0                                                       # 1/1024 forced zero: the error path
        x_sel, s_sel = synth(n_sel, 54, 0, 0), synth(n_sel, 55, 1, 1)
        circuit = [
            {"gadget": pg.OP_ADD_INPUT, "n": n_rc, "group": 0, "values": x_rc},
            {"gadget": pg.OP_RANGE_CHECK, "num_bits": 65, "n": n_rc, "group": 0, "min": mn, "max": mx, "witness": 0},
            {"gadget": pg.OP_ADD_INPUT, "n": n_mb, "group": 1, "values": x_mb},
            {"gadget": pg.OP_MAX_BOUND, "num_bits": 253, "n": n_mb, "group": 1, "max": mx252, "witness": 2},
            {"gadget": pg.OP_ADD_INPUT, "n": n_nz, "group": 2, "values": x_nz},
            {"gadget": pg.OP_IS_NON_ZERO, "n": n_nz, "group": 2, "var": 4, "assigned": x_nz},
            {"gadget": pg.OP_ADD_INPUT, "n": n_sel, "group": 3, "values": x_sel},
            {"gadget": pg.OP_ADD_INPUT, "n": n_sel, "group": 3, "values": s_sel},
            {"gadget": pg.OP_SELECT_ONE, "n": n_sel, "group": 3, "x": 6, "select": 7},
            {"gadget": pg.OP_SELECT_ZERO, "n": n_sel, "group": 3, "x": 8, "select": 7},
        ]
        total_rows = sum(cc["n"] * pg.op_shape(cc["gadget"], cc.get("num_bits", 0))[0] for cc in circuit)
        n_zero = n_nz // 1024
        for policy_name, policy in (("rows", pg.SHARD_ROWS), ("even", pg.SHARD_EVEN)):
            mine = sharding.plan_of(circuit, world, policy)[rank]
            my_rows = sum((s.inst_hi - s.inst_lo) * pg.op_shape(cc["gadget"], cc.get("num_bits", 0))[0] for s, cc in zip(mine, circuit))

            def step():
                c.reset()
                _, n_err = sharding.run_circuit(c, circuit, mine)
                return c.check_sharded(mine, n_err)                                # local check + NCCL all-reduce, every step

            for _ in range(max(args.warmup, 1)):
                v = step()
            assert v[0] == n_zero and v[2] == n_zero, v                            # every forced zero: one error + one unsatisfied row
            first_expected = 3 + n_rc * 271 + n_mb * 511 + 1023 * 3 + 2
            assert v[1] == first_expected, (v, first_expected)                     # ... numbered as in the sequential composer
            c.timing(reset=True)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
            rows_t = torch.tensor([my_rows], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                lo_t, hi_t = rows_t.clone(), rows_t.clone()
                dist.all_reduce(lo_t, op=dist.ReduceOp.MIN); dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
            else:
                lo_t = hi_t = rows_t
            tim = c.timing(reset=True)
            if rank == 0:
                print(json.dumps({"config": f"C5: mixed circuit (range_check k=65 / max_bound k=253 / is_non_zero with 1/1024 zeros / select_one+select_zero), "
                                            f"{total_rows} rows, sharded at op boundaries by prefix-summed row counts",
                                  "n_gpus": world, "scaling": "strong", "policy": policy_name, "check_mode": mode_name, "ms_per_step": float(t[0]),
                                  "gate_evals_per_s": total_rows / (float(t[0]) * 1e-3), "rows_per_rank_min_max": [int(lo_t[0]), int(hi_t[0])],
                                  "verdict": {"n_unsat": v[0], "first_bad_row": v[1], "n_err": v[2]},
                                  "collective": "pg_check_sharded: ncclAllReduce(sum, min) of the verdict inside every step",
                                  "rank0_kernel_ms": {k: tim[k] / args.steps for k in ("check_ms", "witness_ms", "other_ms")}}), flush=True)

        # ---- gather of witness shards: the Variables of one range_check call (653 per instance, 32 B each) from every rank to every rank
        n_g = 1 << args.gather_log2n
        c.reset()
        w = c.add_input(x_rc[:n_g] if n_g <= n_rc else synth(n_g, 56, 2, 64))
        pg.range_check(c, mn, mx, w)
        per_rank = n_g * 653
        dst = torch.empty((world * per_rank, 4), dtype=torch.int64, device=dev)
        for _ in range(2):
            assert c.gather_variables(1, dst) == world * per_rank
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        c.timing(reset=True)
        e0.record(stream)
        for _ in range(reps):
            c.gather_variables(1, dst)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / reps
        tim = c.timing(reset=True)
        res = torch.empty((world * n_g, 4), dtype=torch.int64, device=dev)
        y = pg.Variables(c, w.col + 1, n_g)
        tg = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        if rank == 0 and mode == pg.CHECK_GENERIC:
            recv = (world - 1) * per_rank * 32
            print(json.dumps({"config": f"gather of witness shards: pg_gather_variables of a range_check call, 2^{args.gather_log2n} instances x 653 Variables per rank",
                              "n_gpus": world, "bytes_per_rank_shard": per_rank * 32, "bytes_received_per_gpu": recv, "ms": float(tg[0]),
                              "ms_includes": "expansion of the packed table into Variable order (ReadVarsBody) + the count exchange + the NCCL all-gather",
                              "nccl_ms": tim["other_ms"] / reps, "nccl_ms_note": "all kernels of class 'other' on rank 0 (expansion + collective)",
                              "recv_GB_per_s_per_gpu": recv / (float(tg[0]) * 1e-3) / 1e9 if world > 1 else None}), flush=True)
        c.comm_destroy()
        c.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
