"""Throughput of the evaluation-domain kernels: pg_fft on device-resident vectors and pg_wire_polynomials.
One JSON line per size; CUDA events on the engine's stream; inputs resident in HBM."""
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


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def main():
    c = pg.StandardComposer(device=0, stream=stream.cuda_stream)
    sizes = [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24, 26]
    for log_n in sizes:
        n = 1 << log_n
        x = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 7, 0, 0, x)
        ms = timed(lambda: c.fft(x), steps=5 if log_n <= 24 else 3)
        bf = log_n * (n // 2)
        print(json.dumps({"op": "pg_fft (forward, in place)", "log_n": log_n, "ms": ms, "butterflies_per_s": bf / (ms * 1e-3),
                          "scalars_per_s": n / (ms * 1e-3), "GB_per_s_one_pass_equiv": 2 * n * 32 / (ms * 1e-3) / 1e9}), flush=True)
        ms = timed(lambda: c.fft(x, inverse=True), steps=3)
        print(json.dumps({"op": "pg_fft (inverse, in place)", "log_n": log_n, "ms": ms, "butterflies_per_s": bf / (ms * 1e-3)}), flush=True)
        del x
    # wire polynomials of 2^16 range_check instances (17.8 M rows -> domain 2^25)
    R2 = np.array([[0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]], dtype=np.uint64)
    raw = np.array([[0, 0, 0, 0], [0, 1, 0, 0]], dtype=np.uint64)
    b = c.fr_op(0, raw, np.repeat(R2, 2, axis=0)); mn, mx = b[0:1].copy(), b[1:2].copy()
    n = 1 << 16
    wit = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit); pg.range_check(c, mn, mx, w)
    k = c.domain_log_size()
    polys = torch.empty((4, 1 << k, 4), dtype=torch.int64, device=dev)
    ms = timed(lambda: c.wire_polynomials(out=polys), steps=3, warmup=1)
    print(json.dumps({"op": "pg_wire_polynomials (4 columns: materialise + pad + ifft)", "rows": c.circuit_size(), "log_n": k, "ms": ms,
                      "rows_per_s": c.circuit_size() / (ms * 1e-3), "butterflies_per_s": 4 * k * (1 << (k - 1)) / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
