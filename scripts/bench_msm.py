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
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sizes = [int(a) for a in args] or ([] if "--round1" in sys.argv else [16, 18, 20])
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
    if "--round1" in sys.argv:
        # The prover's first round for a batch of 2^16 range_check instances (17.8 M rows, domain 2^25): gadgets + verdict,
        # wire polynomials (4 inverse FFTs), and the four commitments (4 MSMs of 2^25 terms) against a device-resident SRS.
        import time
        raw = np.array([[0, 0, 0, 0], [0, 1, 0, 0]], dtype=np.uint64)
        b = c.fr_op(0, raw, np.repeat(R2, 2, axis=0)); mn, mx = b[0:1].copy(), b[1:2].copy()
        n = 1 << 16
        wit = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 2, 2, 64, wit)
        k = 25
        srs = torch.empty((1 << k, 12), dtype=torch.int64, device=dev)
        t0 = time.perf_counter(); c.srs_powers(beta, 1 << k, out=srs); torch.cuda.synchronize(dev); srs_s = time.perf_counter() - t0
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for rep in range(2):
            c.reset()
            torch.cuda.synchronize(dev); e[0].record(stream)
            w = c.add_input(wit); pg.range_check(c, mn, mx, w)
            bad, _ = c.check_circuit_satisfied(); assert bad == 0 and c.domain_log_size() == k
            e[1].record(stream)
            polys = torch.empty((4, 1 << k, 4), dtype=torch.int64, device=dev)
            c.wire_polynomials(out=polys)
            e[2].record(stream)
            coms = [c.msm(srs, polys[col]) for col in range(4)]
            e[3].record(stream); torch.cuda.synchronize(dev)
        again = c.commit_wire_polynomials(srs)
        assert all((again[col] == coms[col]).all() for col in range(4))
        # the same four commitments from the wire VALUES against the Lagrange-basis form of the same SRS: no FFT, mostly-empty windows
        lag = torch.empty((1 << k, 12), dtype=torch.int64, device=dev)
        t0 = time.perf_counter(); c.srs_lagrange(beta, k, out=lag); torch.cuda.synchronize(dev); lag_s = time.perf_counter() - t0
        f = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(2):
            torch.cuda.synchronize(dev); f[0].record(stream)
            ev = c.commit_wire_evaluations(lag)
            f[1].record(stream); torch.cuda.synchronize(dev)
        assert all((ev[col] == coms[col]).all() for col in range(4))
        print(json.dumps({"op": "prover round 1: 2^16 range_check instances -> verdict -> wire polynomials -> 4 KZG commitments",
                          "rows": c.circuit_size(), "log_domain": k, "gadgets_and_check_ms": e[0].elapsed_time(e[1]),
                          "wire_polynomials_ms": e[1].elapsed_time(e[2]), "four_commitments_ms": e[2].elapsed_time(e[3]),
                          "total_ms": e[0].elapsed_time(e[3]), "srs_setup_s": srs_s,
                          "four_commitments_from_wire_values_lagrange_srs_ms": f[0].elapsed_time(f[1]), "lagrange_srs_setup_s": lag_s}), flush=True)


if __name__ == "__main__":
    main()
