"""Throughput of dusk-plonk's native range gate: pg_range_gate_batch + pg_check.
One JSON line per width; witnesses resident in HBM, CUDA events on the engine's stream, verdicts asserted.
usage: python scripts/bench_range_gate.py [log2n] [num_bits ...]"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import plonk_gadgets_b200 as pg

SEED = 0x706C6F6E6B5F6732
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def main():
    args = [int(a) for a in sys.argv[1:]]
    log2n = args[0] if args else 24
    widths = args[1:] or [16, 64, 128, 254]
    n = 1 << log2n
    shape = int(os.environ.get("PG_CHECK_SHAPE", "0"))           # launch shape of the gate-check kernels (tuning runs)
    mode = pg.CHECK_SPARSE if os.environ.get("PG_CHECK_MODE") == "sparse" else pg.CHECK_GENERIC
    c = pg.StandardComposer(device=0, timing=True, stream=stream.cuda_stream, check_shape=shape, check_mode=mode)
    wit = torch.empty((n, 4), dtype=torch.int64, device=dev)
    for bits in widths:
        c.synth(SEED, 61, 1, bits, wit)
        gates = (bits + 7) // 8

        def step():
            c.reset(); w = c.add_input(wit); c.range_gate(w, bits)
            bad, _ = c.check_circuit_satisfied(); assert bad == 0
        for _ in range(2):
            step()
        c.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 3
        torch.cuda.synchronize(dev); e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        tim = c.timing(reset=True)
        chk, wit_ms = tim["check_ms"] / steps, tim["witness_ms"] / steps
        rows = (gates + 2) * n
        wide = (gates * 752 + 416) * n                                  # executed wide multiplier instructions
        print(json.dumps({
            "op": "pg_range_gate_batch + pg_check", "log2n": log2n, "num_bits": bits, "rows_per_witness": gates + 2, "vars_per_witness": bits // 2,
            "ms_per_step": ms, "check_ms": chk, "witness_ms": wit_ms, "witnesses_per_s": n / (ms * 1e-3), "gate_evals_per_s": rows / (ms * 1e-3),
            "range_rows_per_s_in_kernel": gates * n / (chk * 1e-3), "executed_wide_products_per_s_in_kernel": wide / (chk * 1e-3),
            "witness_GB_per_s_written": n * (bits // 2) * 32 / (wit_ms * 1e-3) / 1e9,
            "same_bound_with_range_check_rows": 4 * (bits + 1) + 11 if bits <= 252 else None}), flush=True)


if __name__ == "__main__":
    main()
