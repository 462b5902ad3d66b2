"""Throughput of the commitment kernels (second half): pg_msm on device-resident points and scalars.
One JSON line per size; CUDA events on the engine's stream."""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import plonk_gadgets_b200 as pg

SEED = 0x706C6F6E6B5F6732
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def main():
    c = pg.StandardComposer(device=0, timing=True, stream=stream.cuda_stream)
    sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 20]
    R2 = np.array([[0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]], dtype=np.uint64)
    beta = c.fr_op(0, np.array([[0x1234567, 0x89abcdef, 0x42, 0]], dtype=np.uint64), R2)[0]
    for log_n in sizes:
        n = 1 << log_n
        srs = torch.empty((n, 12), dtype=torch.int64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev); e0.record(stream)
        c.srs_powers(beta, n, out=srs)
        e1.record(stream); torch.cuda.synchronize(dev)
        srs_ms = e0.elapsed_time(e1)
        sc = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 9, 0, 0, sc)
        c.msm(srs, sc)
        c.timing(reset=True)
        steps = 3
        torch.cuda.synchronize(dev); e0.record(stream)
        for _ in range(steps):
            c.msm(srs, sc)
        e1.record(stream); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"op": "pg_msm (G1, device-resident points and scalars)", "log_n": log_n, "ms": ms, "points_per_s": n / (ms * 1e-3),
                          "srs_powers_ms": srs_ms, "fixed_base_muls_per_s": n / (srs_ms * 1e-3)}), flush=True)
        del srs, sc


if __name__ == "__main__":
    main()
