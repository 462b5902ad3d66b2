"""Bandwidth of the gather of witness shards alone (pg_gather_variables: the 653 Variables per instance of a range_check call, from
every rank to every rank, NCCL over NVLink behind the C ABI).  One process per GPU:
    torchrun --nproc-per-node N scripts/bench_gather.py [log2 instances per rank ...]
One JSON line per size (rank 0).  PG_GATHER_BCAST=1 forces the per-rank broadcasts instead of ncclAllGather."""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import plonk_gadgets_b200 as pg

SEED = 0x706C6F6E6B5F6732
R2 = np.array([[0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]], dtype=np.uint64)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [14, 16]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    uid = torch.zeros(pg.api.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.frombuffer(bytearray(pg.comm_unique_id()), dtype=torch.uint8).to(dev)
    if world > 1:
        dist.broadcast(uid, 0)
    c = pg.StandardComposer(device=local, timing=True, stream=stream.cuda_stream)
    c.comm_init(bytes(uid.cpu().numpy()), rank, world)
    b = c.fr_op(0, np.array([[0, 0, 0, 0], [0, 1, 0, 0]], dtype=np.uint64), np.repeat(R2, 2, axis=0))     # 0 and 2^64 in Montgomery form
    mn, mx = b[0:1].copy(), b[1:2].copy()
    for log2n in sizes:
        n = 1 << log2n
        x = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 56 + rank, 2, 64, x)
        c.reset()
        w = c.add_input(x)
        pg.range_check(c, mn, mx, w)
        per_rank = n * 653
        dst = torch.empty((world * per_rank, 4), dtype=torch.int64, device=dev)
        for _ in range(2):
            assert c.gather_variables(1, dst) == world * per_rank
        # rank r's shard must sit at r * per_rank: its first Variable is the first bit of its own witness
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            c.gather_variables(1, dst)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        own = c.variables(c.num_variables() - per_rank, 4)                      # this rank's first Variables of the call, read back directly
        assert (dst[rank * per_rank: rank * per_rank + 4].cpu().numpy().view(np.uint64) == own).all()
        if rank == 0:
            recv = (world - 1) * per_rank * 32
            print(json.dumps({"config": f"gather of witness shards: pg_gather_variables of a range_check call, 2^{log2n} instances x 653 Variables per rank",
                              "n_gpus": world, "path": "per-rank ncclBroadcast" if os.environ.get("PG_GATHER_BCAST") == "1" else "ncclAllGather (equal shards)",
                              "bytes_per_rank_shard": per_rank * 32, "bytes_received_per_gpu": recv, "ms": float(t[0]),
                              "ms_includes": "expansion of the packed table into Variable order (ReadVarsBody) + the count exchange + the NCCL collective",
                              "recv_GB_per_s_per_gpu": recv / (float(t[0]) * 1e-3) / 1e9 if world > 1 else None}), flush=True)
        del dst, x
    c.comm_destroy()
    c.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
