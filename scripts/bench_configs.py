"""Throughput of the other BASELINE.json configurations (C2-C5 of SURVEY.md section 8) and of the read-back kernels.
One JSON line per configuration; inputs resident in HBM, CUDA events on the engine's stream, verdicts asserted."""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import plonk_gadgets_b200 as pg

SEED = 0x706C6F6E6B5F6732
R2 = np.array([[0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]], dtype=np.uint64)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)      # non-default stream shared by torch (events, tensor ops) and the engine's kernels
torch.cuda.set_stream(stream)


def to_mont(c, canonical_ints):
    raw = np.array([[(v >> (64 * k)) & (2 ** 64 - 1) for k in range(4)] for v in canonical_ints], dtype=np.uint64)
    return c.fr_op(0, raw, np.repeat(R2, len(canonical_ints), axis=0))


def timed(c, fn, steps=3, warmup=2):
    for _ in range(warmup):
        out = fn()
    c.timing(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    tim = c.timing(reset=True)
    return e0.elapsed_time(e1) / steps, {k: (v / steps if k.endswith("_ms") else v) for k, v in tim.items()}, out


def emit(name, rows, ms, tim, extra=None):
    line = {"config": name, "rows_per_step": rows, "ms_per_step": ms, "gate_evals_per_s": rows / (ms * 1e-3),
            "kernel_ms": {k: tim[k] for k in ("check_ms", "witness_ms", "other_ms")}}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    fused = "--fused" in sys.argv                                  # structure-aware + PG_F_FUSED_CHECK: rows evaluated by the gadgets' own kernels
    mode = pg.CHECK_SPARSE if "--sparse" in sys.argv or fused else pg.CHECK_GENERIC
    c = pg.StandardComposer(device=0, check_mode=mode, timing=True, stream=stream.cuda_stream, fused_check=fused)
    zero_2p64 = to_mont(c, [0, 2 ** 64])
    mn, mx = zero_2p64[0:1].copy(), zero_2p64[1:2].copy()

    # ---- C1: ONE range_check instance (latency of the whole path for a single gadget call, host overheads included)
    one = torch.empty((1, 4), dtype=torch.int64, device=dev); c.synth(SEED, 1, 1, 64, one)

    def c1():
        c.reset(); w = c.add_input(one); y = pg.range_check(c, mn, mx, w)
        bad, _ = c.check_circuit_satisfied(); assert bad == 0
        return y.values()
    import time as _t
    for _ in range(20): c1()
    t0 = _t.perf_counter()
    for _ in range(200): c1()
    lat = (_t.perf_counter() - t0) / 200
    print(json.dumps({"config": "C1: single range_check of one 64-bit witness (reset + allocate + gadget + verdict + result read-back)",
                      "latency_us": lat * 1e6, "rows": 271}), flush=True)

    # ---- C2: 2^20 range_check, 64-bit bound
    n = 1 << 20
    wit = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 2, 2, 64, wit)

    def c2():
        c.reset(); w = c.add_input(wit); pg.range_check(c, mn, mx, w)
        bad, _ = c.check_circuit_satisfied(); assert bad == 0
    ms, tim, _ = timed(c, c2)
    emit("C2: 2^20 range_check, 64-bit bound (k=65)", 271 * n, ms, tim)

    # ---- C2 with per-instance random bounds of one width: max_i - 1 in [2^63, 2^64) (k = 65 for every instance),
    #      min_i uniform below 2^63 <= max_i; the two V rows' q_c are per-instance parameters
    mx2 = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 21, 3, 64, mx2)
    mn2 = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 22, 1, 63, mn2)

    def c2b():
        c.reset(); w = c.add_input(wit); pg.range_check(c, mn2, mx2, w)
        bad, _ = c.check_circuit_satisfied(); assert bad == 0
    ms, tim, _ = timed(c, c2b)
    emit("C2b: 2^20 range_check, per-instance random 64-bit bounds (k=65; max_i - 1 in [2^63, 2^64), min_i < 2^63)", 271 * n, ms, tim)
    del mx2, mn2

    # ---- C3: 2^22 max_bound with 252-bit per-instance bounds (k=253)
    n = 1 << 22
    mx3 = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 31, 3, 252, mx3)
    wit3 = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 32, 2, 250, wit3)

    def c3():
        c.reset(); w = c.add_input(wit3); _, k = pg.max_bound(c, mx3, w); assert k == 253
        bad, _ = c.check_circuit_satisfied(); assert bad == 0
    ms, tim, _ = timed(c, c3, steps=2, warmup=1)
    emit("C3: 2^22 max_bound, per-instance 252-bit bounds (k=253, 511 rows each)", 511 * n, ms, tim)
    del mx3, wit3

    # ---- C4: 2^24 x (is_non_zero + maybe_equal), as SURVEY.md 8d specifies it: value uniform Fr with 1/1024 forced zero (the
    #      NonExistingInverse path: every Result kept, pg_is_non_zero_batch_flags), value_assigned = value except 1/1024 mismatches
    #      (their assert_equal row is unsatisfied); maybe_equal with 50 % a = b
    n = 1 << 24
    a = torch.empty((n, 4), dtype=torch.int64, device=dev); b = torch.empty_like(a)
    c.synth(SEED, 41, 0, 0, a); c.synth(SEED, 42, 0, 0, b); b[0::2] = a[0::2]
    assigned = a.clone()
    assigned[1023::1024] = 0                                 # forced zeros: Err(NonExistingInverse), row var*inv - 1 unsatisfied (and var != 0: assert_equal too)
    assigned[511::1024] = b[511::1024]                       # mismatches (odd index: b != a): assert_equal and var*inv - 1 unsatisfied
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    n_zero, n_mis = n // 1024, n // 1024

    def c4():
        c.reset(); va = c.add_input(a); vb = c.add_input(b)
        pg.maybe_equal(c, va, vb)
        pg.is_non_zero_flags(c, va, assigned, pg.NZ_UNIFORM, flags_out=flags)
        assert c.last_n_err == n_zero
        bad, first = c.check_circuit_satisfied()
        assert bad == 2 * n_zero + 2 * n_mis, (bad, n_zero, n_mis)      # assert_equal AND var*inv - 1 fail in both kinds of instance
        return first
    ms, tim, first = timed(c, c4)
    assert int(flags.sum().item()) == n_zero and bool(flags[1023::1024].all())
    assert first == 3 + 3 * n + 3 * 511                      # the assert_equal row of the first mismatching instance (rows: 3 fresh, 3n maybe_equal)
    emit("C4: 2^24 x (is_non_zero + maybe_equal): 2^25 inversions, 6 rows per pair; 1/1024 zeros (per-instance NonExistingInverse flags) and 1/1024 mismatches", 6 * n, ms, tim,
         {"inversions_per_s": 2 * n / (ms * 1e-3), "n_err": n_zero, "n_unsat": 2 * n_zero + 2 * n_mis})
    del a, b, assigned, flags

    # ---- C5: mixed circuit, 2^26 rows: range_check k=65 / max_bound k=253 / is_non_zero / select_one+select_zero, a quarter each
    q = 1 << 24
    n_rc, n_mb, n_nz, n_sel = q // 271, q // 511, q // 3, q // 5
    x_rc = torch.empty((n_rc, 4), dtype=torch.int64, device=dev); c.synth(SEED, 51, 2, 64, x_rc)
    x_mb = torch.empty((n_mb, 4), dtype=torch.int64, device=dev); c.synth(SEED, 52, 2, 250, x_mb)
    x_nz = torch.empty((n_nz, 4), dtype=torch.int64, device=dev); c.synth(SEED, 53, 0, 0, x_nz)
    x_sel = torch.empty((n_sel, 4), dtype=torch.int64, device=dev); c.synth(SEED, 54, 0, 0, x_sel)
    s_sel = torch.empty((n_sel, 4), dtype=torch.int64, device=dev); c.synth(SEED, 55, 1, 1, s_sel)   # 1-bit selectors
    mx252 = to_mont(c, [2 ** 252])

    def c5():
        c.reset()
        w = c.add_input(x_rc); pg.range_check(c, mn, mx, w)
        w = c.add_input(x_mb); pg.max_bound(c, mx252, w)
        w = c.add_input(x_nz); pg.is_non_zero(c, w, x_nz)
        x = c.add_input(x_sel); s = c.add_input(s_sel)
        y = pg.conditionally_select_one(c, x, s); pg.conditionally_select_zero(c, y, s)
        bad, _ = c.check_circuit_satisfied(); assert bad == 0
        return c.circuit_size()
    ms, tim, rows = timed(c, c5)
    emit("C5: mixed circuit (range_check k=65 / max_bound k=253 / is_non_zero / select_one+select_zero), ~2^26 rows", rows - 3, ms, tim)

    # ---- native range gate: 2^24 witnesses, 64 bits each: 10 rows / 32 accumulators per instance, 8 of the rows with
    #      q_range = 1 (twelve multiplications each: four D(f) = f(f-1)(f-2)(f-3) terms)
    n = 1 << 24
    x_rg = torch.empty((n, 4), dtype=torch.int64, device=dev); c.synth(SEED, 61, 1, 64, x_rg)

    def rg():
        c.reset(); w = c.add_input(x_rg); c.range_gate(w, 64)
        bad, _ = c.check_circuit_satisfied(); assert bad == 0
    ms, tim, _ = timed(c, rg)
    emit("range_gate: 2^24 witnesses x 64 bits (dusk-plonk's native quad-accumulator gate; 10 rows, 32 variables per instance)", 10 * n, ms, tim,
         {"witnesses_per_s": n / (ms * 1e-3), "range_rows_per_s": 8 * n / (tim["check_ms"] * 1e-3) if tim["check_ms"] else None,
          "table_GB": n * 33 * 32 / 1e9})
    del x_rg
    c.reset()
    w = c.add_input(x_rc); pg.range_check(c, mn, mx, w)          # the read-back lines below run on a composer of arithmetic rows again
    w = c.add_input(x_mb); pg.max_bound(c, mx252, w)

    # ---- read-back kernels: materialise rows / variables of the last composer to device buffers (reference representation)
    cnt = 1 << 22
    w_idx = torch.empty((4, cnt), dtype=torch.int64, device=dev)
    w_val = torch.empty((4, cnt, 4), dtype=torch.int64, device=dev)
    sel = torch.empty((6, cnt, 4), dtype=torch.int64, device=dev)
    pi = torch.empty((cnt, 4), dtype=torch.int64, device=dev)
    import ctypes as C

    def mat():
        c._ok(c._L.pg_materialize_rows(c._ctx, 3, cnt, C.c_void_p(w_idx.data_ptr()), C.c_void_p(w_val.data_ptr()), C.c_void_p(sel.data_ptr()),
                                       C.c_void_p(pi.data_ptr()), 1), "pg_materialize_rows")
    ms, tim, _ = timed(c, mat)
    emit("materialize 2^22 rows (4 wire ids + 4 wire values + 6 selectors + PI = 384 B/row) to device buffers", cnt, ms, tim,
         {"GB_per_s_written": cnt * 384 / (ms * 1e-3) / 1e9})
    sig = torch.empty((4, cnt), dtype=torch.int64, device=dev)

    def perm():
        c._ok(c._L.pg_permutation(c._ctx, 3, cnt, C.c_void_p(sig.data_ptr()), 1), "pg_permutation")
    ms, tim, _ = timed(c, perm)
    emit("permutation map (copy-constraint cycle successors) of 2^22 rows, 32 B/row written", cnt, ms, tim, {"GB_per_s_written": cnt * 32 / (ms * 1e-3) / 1e9})
    vars_buf = torch.empty((cnt, 4), dtype=torch.int64, device=dev)

    def rv():
        c._ok(c._L.pg_read_variables(c._ctx, 5, cnt, C.c_void_p(vars_buf.data_ptr()), 1), "pg_read_variables")
    ms, tim, _ = timed(c, rv)
    emit("read 2^22 variables (Variable order, 32 B each) to a device buffer", cnt, ms, tim, {"GB_per_s_written": cnt * 32 / (ms * 1e-3) / 1e9})


if __name__ == "__main__":
    main()
