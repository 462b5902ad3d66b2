"""Parity tests proper: the CUDA path (libpg_b200.so, through the C ABI) against the CPU oracle and the golden vectors.
Bit-exact everywhere (integer arithmetic mod q); sizes the oracle finishes in seconds, plus size-independent properties at
the BASELINE.json sizes.  Run on the B200 box with `pytest -m gpu`."""
import ctypes as C
import os
import random

import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from tests.engine_runner import run_engine
from tests.programs import Q, SEED, hx, run_oracle, synth_wide

pytestmark = pytest.mark.gpu

R = (1 << 256) % Q


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


def gpu_composer(**kw):
    return pg.StandardComposer(device=0, **kw)


def test_cuda_library_is_what_runs():
    """No fallback: the composer is backed by the in-tree CUDA library and a real device context."""
    c = gpu_composer()
    with open("/proc/self/maps") as f:
        assert "libpg_b200.so" in f.read()
    assert c.circuit_size() == 3 and c.num_variables() == 5
    assert c.check_circuit_satisfied() == (0, None)


def test_fr_kernels_vs_oracle(oracle):
    rng = random.Random(11)
    edge = [0, 1, 2, Q - 1, Q - 2, R, Q - R, 2 ** 32 - 1, 2 ** 64 - 1, (1 << 254) + 12345, Q >> 1, (0xffffffff << 224) % Q, Q - 2 ** 224]
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(Q), rng.randrange(Q)) for _ in range(20000)]
    A = oracle.from_ints([p[0] for p in pairs]); B = oracle.from_ints([p[1] for p in pairs])
    c = gpu_composer()
    exp = {0: lambda a, b: a * b % Q, 6: lambda a, b: a * b % Q, 1: lambda a, b: (a + b) % Q, 2: lambda a, b: (a - b) % Q}
    for op, f in exp.items():
        got = oracle.to_ints(c.fr_op(op, A, B))
        assert got == [f(a, b) for a, b in pairs], f"fr op {op}"
    assert oracle.to_ints(c.fr_op(3, A)) == [(-a) % Q for a, _ in pairs]
    # from_mont: raw limbs of the result are the canonical integer
    fm = c.fr_op(5, A)
    assert [int.from_bytes(fm[i].tobytes(), "little") for i in range(len(pairs))] == [a for a, _ in pairs]


@pytest.mark.parametrize("n", [1, 31, 32, 255, 256, 257, 1000, 4096 + 17, 620_003])
def test_block_batch_inversion(oracle, n):
    """Montgomery's trick across a thread block: zeros anywhere (including whole warps / whole blocks), ragged tails; the last
    size is past the point where the one-wave grid gives every thread more than the minimum of 8 elements."""
    rng = random.Random(n)
    vals = [rng.randrange(Q) for _ in range(n)]
    for i in range(0, n, 7):
        vals[i] = 0
    if n >= 64:
        for i in range(32, 64):
            vals[i] = 0                       # a whole warp of zeros
    if n >= 1000:
        for i in range(512, 768):
            vals[i] = 0                       # a whole block of zeros
    c = gpu_composer()
    got = oracle.to_ints(c.fr_op(4, oracle.from_ints(vals)))
    assert got == [0 if v == 0 else pow(v, -1, Q) for v in vals]


def test_golden_programs(golden, oracle):
    """Every golden program (reference KATs + batched programs): counts, verdict, error behaviour, returned values and the
    digest of the complete composer state."""
    for name, spec in golden.items():
        exp = spec["expected"]
        snap = run_engine(spec["program"], gpu_composer, oracle)
        assert (snap.n_rows, snap.n_vars) == (exp["n_rows"], exp["n_vars"]), name
        assert snap.unsat == exp["unsat"], name
        assert (list(snap.error) if snap.error else None) == exp["error"], name
        assert snap.digest() == exp["digest"], name
        for k, vals in exp["results"].items():
            assert [hx(v) for v in snap.results(int(k))] == vals, (name, k)
        if "satisfied" in spec:
            assert (len(snap.unsat) == 0) == spec["satisfied"], name


@pytest.mark.parametrize("mode", [pg.CHECK_GENERIC, pg.CHECK_SPARSE])
def test_check_modes_agree(golden, oracle, mode):
    for name in ("batch_mixed_circuit", "batch_max_bound_k8_claims", "kat_range_check_1_wrongclaim", "kat_is_non_zero_mismatch"):
        spec = golden[name]
        snap = run_engine(spec["program"], lambda: gpu_composer(check_mode=mode), oracle)
        assert snap.unsat == spec["expected"]["unsat"], name


@pytest.mark.parametrize("bits", [1, 2, 31, 32, 33, 64, 65, 127, 128, 200, 252, 253])
def test_range_check_vs_oracle(oracle, bits):
    """Random and boundary witnesses (min, max-1, max, min-1, 0, q-1, 2^k ...) at many bit widths, against the C oracle run on
    the same program: full composer state equal."""
    r = synth_wide(50 + bits, 300)
    mx = ((r[0] % 2 ** (bits - 1)) | 2 ** (bits - 1)) + 1 if bits > 1 else 2
    mn = r[1] % mx
    wit = [mn, mx - 1, mx, (mn - 1) % Q, 0, Q - 1, 2 ** bits, 2 ** bits - 1, 1] + [x % mx for x in r[2:150]] + r[150:]
    prog = [dict(op="add_input", values=[hx(x) for x in wit]), dict(op="range_check", min=hx(mn), max=hx(mx), witness=0),
            dict(op="max_bound", max=hx(mx), witness=0)]
    so = run_oracle(prog)
    se = run_engine(prog, gpu_composer, oracle)
    assert se.digest() == so.digest()
    assert se.unsat == so.unsat == []
    assert se.results(1) == so.results(1) and se.results(2) == so.results(2)


def test_per_instance_bounds_and_mixed_bits(oracle):
    r = synth_wide(77, 400)
    mx = [((x % 2 ** 63) | 2 ** 63) + 1 for x in r[:100]]
    mn = [r[100 + i] % mx[i] for i in range(100)]
    wit = [(r[200 + i] % (2 ** 65)) if i % 3 else mn[i] for i in range(100)]
    prog = [dict(op="add_input", values=[hx(x) for x in wit]),
            dict(op="range_check", min=[hx(x) for x in mn], max=[hx(x) for x in mx], witness=0)]
    so, se = run_oracle(prog), run_engine(prog, gpu_composer, oracle)
    assert se.digest() == so.digest() and se.unsat == []
    c = gpu_composer()
    w = c.add_input(oracle.from_ints([1, 2, 3]))
    with pytest.raises(pg.EngineError) as e:
        pg.range_check(c, oracle.from_ints([0, 0, 0]), oracle.from_ints([2 ** 10, 2 ** 10, 2 ** 20]), w)
    assert e.value.code == -4 and c.circuit_size() == 3 and c.check_circuit_satisfied() == (0, None)


def test_scalar_gadgets_vs_oracle(oracle):
    r = synth_wide(5, 3000)
    n = 1000
    a = r[:n]; b = [a[i] if i % 2 == 0 else r[n + i] for i in range(n)]
    sel = [i % 2 for i in range(n)]
    prog = [dict(op="add_input", values=[hx(x) for x in a]), dict(op="add_input", values=[hx(x) for x in b]),
            dict(op="add_input", values=[hx(x) for x in sel]),
            dict(op="maybe_equal", a=0, b=1), dict(op="select_zero", x=0, select=2), dict(op="select_one", y=1, select=3),
            dict(op="is_non_zero", var=0, assigned=[hx(a[i] if i % 5 else r[2 * n + i]) for i in range(n)])]
    so, se = run_oracle(prog), run_engine(prog, gpu_composer, oracle)
    assert se.digest() == so.digest()
    assert se.unsat == so.unsat and len(se.unsat) == 2 * (n // 5)      # mismatching value_assigned: two rows each
    # is_non_zero stops at the first zero like `?`
    vals = r[:700] + [0] + r[700:900] + [0]
    prog = [dict(op="add_input", values=[hx(x) for x in vals]), dict(op="is_non_zero", var=0, assigned=[hx(x) for x in vals])]
    so, se = run_oracle(prog), run_engine(prog, gpu_composer, oracle)
    assert se.error == so.error == (1, "NonExistingInverse", 700)
    assert se.digest() == so.digest()


def test_synth_matches_python_generator(oracle, torch_cuda):
    torch = torch_cuda
    c = gpu_composer()
    n = 513
    buf = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    c.synth(SEED, 3, 0, 0, buf); c.sync()
    assert oracle.to_ints(buf.cpu().numpy().view(np.uint64)) == synth_wide(3, n)
    c.synth(SEED, 3, 1, 64, buf); c.sync()
    assert oracle.to_ints(buf.cpu().numpy().view(np.uint64)) == _low_bits(3, n, 64)
    c.synth(SEED, 3, 3, 64, buf); c.sync()
    assert oracle.to_ints(buf.cpu().numpy().view(np.uint64)) == [v | 2 ** 63 for v in _low_bits(3, n, 64)]


def _low_bits(stream, n, bits):
    from tests.programs import splitmix64
    words = splitmix64(SEED ^ (stream * 0xD1342543DE82EF95 & (2 ** 64 - 1)), 8 * n).reshape(n, 8)
    out = []
    for row in words:
        v = sum(int(row[j]) << (64 * j) for j in range(4))
        out.append(v & (2 ** bits - 1))
    return out


def test_fault_injection_rows(oracle):
    c = gpu_composer()
    w = c.add_input(oracle.from_ints([12345, 2 ** 70, Q - 5]))
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    rows = c.rows()
    assert c.check_rows(rows["w_val"], rows["sel"], rows["pi"]) == (0, None)
    rng = random.Random(2)
    for _ in range(20):
        r_, col = rng.randrange(3, c.circuit_size()), rng.randrange(3)
        bad = rows["w_val"].copy()
        bad[col, r_] = oracle.from_ints([rng.randrange(1, Q)])[0]
        n_bad, first = c.check_rows(bad, rows["sel"], rows["pi"])
        # a changed wire can only break its own row; selector 0 on that wire leaves the row satisfied
        assert (n_bad, first) in ((1, r_), (0, None))


@pytest.mark.parametrize("log2n", [20, 22])
def test_range_check_c2_properties(oracle, torch_cuda, log2n):
    """BASELINE config C2 (2^20 range_check, 64-bit bound) and 4x larger: verdict 0 unsatisfied rows; even instances (uniform
    u64) are in range, odd ones (uniform Fr) are not; a random sample of instances is compared row by row with the oracle."""
    torch = torch_cuda
    n = 1 << log2n
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit)
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    assert c.circuit_size() == 3 + 271 * n and c.num_variables() == 5 + n + 653 * n
    assert c.check_circuit_satisfied() == (0, None)
    res = y.values()
    one = oracle.from_ints([1])[0]
    assert (res[0::2] == one).all() and (res[1::2] == 0).all()
    # sample: instances -> the oracle's single-instance composer, renumbered
    rng = random.Random(log2n)
    sample = [0, 1, n - 2, n - 1] + [rng.randrange(n) for _ in range(12)]
    wv = w.values()
    for i in sample:
        oc = oracle.Composer()
        ov = oc.add_input_batch(wv[i:i + 1])
        oc.range_check_batch(oracle.from_ints([0]), oracle.from_ints([2 ** 64]), ov)
        o_vars = oc.variables()[6:]                              # the 653 variables of the gadget
        g_vars = c.variables(5 + n + 653 * i, 653)
        assert (o_vars == g_vars).all(), f"variables of instance {i}"
        rows = c.rows(3 + 271 * i, 271, want=("w_idx", "sel"))
        o_w = oc.wires()[:, 3:]; o_sel = oc.selectors()[:6, 3:]
        # oracle numbering: witness = var 5, gadget vars 6.. ; engine: witness = 5+i, gadget vars 5+n+653*i ..
        remap = np.where(o_w == 0, 0, np.where(o_w == 5, 5 + i, o_w - 6 + 5 + n + 653 * i)).astype(np.uint64)
        assert (rows["w_idx"] == remap).all(), f"wires of instance {i}"
        assert (rows["sel"] == o_sel).all(), f"selectors of instance {i}"


def test_chunked_host_input_equals_device_input(oracle, torch_cuda):
    """Host batches of >= 2^20 scalars are copied in chunks on the input stream while RangePre already runs on the chunks that
    have arrived: every variable must equal the run with device-resident input -- through the chunk-aware consumer (range_check
    right after add_input), through consumers that wait for the whole copy (maybe_equal; range_check on the first of two pending
    columns; a read-back), from pinned and from pageable host memory."""
    torch = torch_cuda
    n = (1 << 20) + 777                                   # ragged last chunk
    mn, mx = oracle.from_ints([0]), oracle.from_ints([2 ** 64])
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    c.synth(SEED, 2, 2, 64, wit)
    c.sync()
    y = pg.range_check(c, mn, mx, c.add_input(wit))
    assert c.check_circuit_satisfied() == (0, None)
    ref_y = y.values()
    ref_vars = [c.variables(5 + n + 653 * i, 653).copy() for i in (0, n // 4 - 1, n // 4 + 300, n // 2 + 5, n - 1)]
    pinned = wit.cpu().pin_memory()
    pageable = pinned.numpy().view(np.uint64).copy()
    for host in (pinned, pageable):
        c.reset()
        y = pg.range_check(c, mn, mx, c.add_input(host))          # chunk by chunk behind the copies
        assert c.check_circuit_satisfied() == (0, None)
        assert (y.values() == ref_y).all()
        for i, ref in zip((0, n // 4 - 1, n // 4 + 300, n // 2 + 5, n - 1), ref_vars):
            assert (c.variables(5 + n + 653 * i, 653) == ref).all(), i
    # consumers that are not chunk-aware
    c.reset()
    a = c.add_input(pinned); b = c.add_input(pageable)             # two pending columns
    assert (b.values(n - 3, 3) == pageable[n - 3:]).all()          # read-back waits for the copies
    eq = pg.maybe_equal(c, a, b)
    y = pg.range_check(c, mn, mx, a)
    assert c.check_circuit_satisfied() == (0, None)
    assert (eq.values() == oracle.from_ints([1])[0]).all() and (y.values() == ref_y).all()
    # a write into a column that is still arriving lands after the copy
    c.reset()
    a = c.add_input(pinned)
    c.poke_variable(5 + n - 2, oracle.from_ints([424242])[0])
    assert (a.values(n - 2, 1) == oracle.from_ints([424242])).all() and (a.values(n - 1, 1) == pageable[n - 1:]).all()
    c.reset()
    a = c.add_input(pinned); b = c.add_input(pinned)               # both still in flight when range_check(a) starts
    y = pg.range_check(c, mn, mx, a)
    assert c.check_circuit_satisfied() == (0, None) and (y.values() == ref_y).all()


def test_max_bound_c3_and_scalar_c4_properties(oracle, torch_cuda):
    """C3 shape (max_bound, 252-bit per-instance bounds, k=253) at 2^18 and C4 shape (is_non_zero + maybe_equal) at 2^20."""
    torch = torch_cuda
    n = 1 << 18
    c = gpu_composer()
    mx = torch.empty((n, 4), dtype=torch.int64, device="cuda"); wit = torch.empty_like(mx)
    c.synth(SEED, 31, 3, 252, mx)           # max in [2^251, 2^252): max-1 has 252 bits unless max = 2^251 exactly
    c.synth(SEED, 32, 2, 250, wit)          # even: 250-bit (below every bound), odd: uniform Fr
    w = c.add_input(wit)
    y, k = pg.max_bound(c, mx, w)
    assert k == 253 and c.circuit_size() == 3 + 511 * n
    assert c.check_circuit_satisfied() == (0, None)
    res = y.values(); one = oracle.from_ints([1])[0]
    assert (res[0::2] == one).all()
    frac_in = float((res[1::2] == one).all(axis=1).mean())
    assert 0.2 < frac_in < 0.35              # y = 1 iff (max-1-x mod q) fits k = 253 bits: probability 2^253/q ~ 0.277
    # C4 (2^20 pairs: past the size where the one-wave batch inversion gives a thread more than its minimum of 8 elements)
    c.reset()
    n = 1 << 20
    a = torch.empty((n, 4), dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
    c.synth(SEED, 41, 0, 0, a); c.synth(SEED, 42, 0, 0, b)
    c.sync()                                 # the engine runs on its own stream: finish before torch touches the buffers
    b[0::2] = a[0::2]
    torch.cuda.synchronize()
    va, vb = c.add_input(a), c.add_input(b)
    eq = pg.maybe_equal(c, va, vb)
    pg.is_non_zero(c, va, a)
    assert c.check_circuit_satisfied() == (0, None)
    r = eq.values()
    assert (r[0::2] == one).all() and (r[1::2] == 0).all()
    # the inverses themselves, at both ends of the batch: variables u, z, y of instance i follow the 5 + 2n inputs
    for i0 in (0, n - 64):
        uzy = oracle.to_ints(c.variables(5 + 2 * n + 3 * i0, 3 * 64))
        for j in range(64):
            u, z, y = uzy[3 * j: 3 * j + 3]
            assert (u, z, y) == (0, 0, 1) if (i0 + j) % 2 == 0 else (u != 0 and z * u % Q == 1 and y == 0)


def test_full_size_metric_config(oracle, torch_cuda):
    """The headline configuration: 2^24 range_check instances, 64-bit bound (4.55e9 rows) -- verdict and result pattern."""
    torch = torch_cuda
    free, _total = torch.cuda.mem_get_info()
    n = 1 << 24
    if free < 90 * 2 ** 30:
        pytest.skip("needs ~80 GB of free HBM")
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit)
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    assert c.check_circuit_satisfied() == (0, None)
    res = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    c.read_column_into(y, res); c.sync()
    one = torch.from_numpy(oracle.from_ints([1]).view(np.int64)).cuda()
    assert bool((res[0::2] == one).all()) and bool((res[1::2] == 0).all())
    c.close()


def test_wire_format_gpu(oracle):
    rng = random.Random(3)
    vals = [0, 1, Q - 1, 2 ** 255 % Q] + [rng.randrange(Q) for _ in range(5000)]
    c = gpu_composer()
    m = oracle.from_ints(vals)
    raw = c.to_bytes(m)
    assert [int.from_bytes(raw[i].tobytes(), "little") for i in range(len(vals))] == vals
    back, bad, first = c.from_bytes(raw)
    assert bad == 0 and first is None and (back == m).all()
    bogus = raw.copy()
    bogus[70] = np.frombuffer(Q.to_bytes(32, "little"), dtype=np.uint8)
    bogus[900] = 0xFF
    back, bad, first = c.from_bytes(bogus)
    assert (bad, first) == (2, 70) and (back[70] == 0).all() and (back[900] == 0).all() and (back[71] == m[71]).all()


def test_microbench_sanity():
    """Roofline denominators: 32-bit IMAD near 64/SM/clk, wide products near 32/SM/clk."""
    import torch
    c = gpu_composer()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    lo, wide = c.microbench(0), c.microbench(1)
    assert 40 < lo / sms / 1.965e9 < 70 and 20 < wide / sms / 1.965e9 < 36


def test_permutation_map_gpu(golden, oracle):
    """Copy-constraint cycles (pg_permutation) against the oracle's perm.variable_map on the batched golden programs."""
    from tests.programs import expected_sigma
    for name in ("batch_mixed_circuit", "batch_range_check_k65", "batch_is_non_zero_error_midway", "batch_is_non_zero_maybe_equal", "kat_select_one_sel1"):
        spec = golden[name]
        so, oc = run_oracle(spec["program"], return_composer=True)
        se, c = run_engine(spec["program"], gpu_composer, oracle, return_composer=True)
        assert (c.permutation() == expected_sigma(oc)).all(), name
    # cycle property at scale: sigma is a permutation of all 4*n positions and preserves the variable on the wire
    n = 1 << 14
    c = gpu_composer()
    w = c.add_input(oracle.from_ints(list(range(n))))
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    rows = c.circuit_size()
    sigma = c.permutation()
    assert sigma.shape == (4, rows)
    flat = np.sort(sigma.reshape(-1))
    # positions are row*4 + wire; sigma output is column-major by wire, so compare as sets
    assert (flat == np.arange(4 * rows, dtype=np.uint64)).all()
    w_idx = c.rows(want=("w_idx",))["w_idx"]
    nxt_row, nxt_wire = (sigma // 4).astype(np.int64), (sigma % 4).astype(np.int64)
    assert (w_idx[nxt_wire, nxt_row] == w_idx).all()


def test_materialize_tiled_gpu(oracle):
    """The tiled materialisation kernel (smem transpose + TMA bulk stores) and the simple body on ragged ranges vs the oracle."""
    from tests.test_emu_engine import _materialize_vs_oracle
    _materialize_vs_oracle(gpu_composer, oracle, n=70)
    _materialize_vs_oracle(gpu_composer, oracle, n=1000)


def test_async_result_copy(oracle, torch_cuda):
    """pg_col_read with dst_on_device = 2: pinned host destination filled on the copy stream, valid after pg_sync."""
    torch = torch_cuda
    n = 1 << 16
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit)
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    host = torch.zeros((n, 4), dtype=torch.int64).pin_memory()
    c.read_column_into(y, host, asynchronous=True)
    assert c.check_circuit_satisfied() == (0, None)
    c.sync()
    assert (host.numpy().view(np.uint64) == y.values()).all()


@pytest.mark.parametrize("mode,fused", [(pg.CHECK_GENERIC, False), (pg.CHECK_SPARSE, False), (pg.CHECK_SPARSE, True)])
def test_chunked_range_pipeline_with_async_read(oracle, torch_cuda, mode, fused):
    """Behind a chunked input copy the range gadgets run decomposition / inversion / results chunk by chunk, and an asynchronous read
    of the whole result column is issued per chunk on the copy stream (it overlaps with the later chunks' kernels): results, the
    table and the verdict must equal the device-input run; range_check and max_bound, per-instance bounds too."""
    torch = torch_cuda
    n = (1 << 20) + 4099                                  # ragged last chunk
    c = gpu_composer(check_mode=mode, fused_check=fused)
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 2, 2, 64, wit); c.sync()
    mxs = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 77, 3, 64, mxs); c.sync()     # 64-bit bounds with the top bit set
    mn, mx = oracle.from_ints([0]), oracle.from_ints([2 ** 64])
    pinned = wit.cpu().pin_memory()
    for gadget in ("range_check", "max_bound_per_instance"):
        def build(src):
            c.reset()
            w = c.add_input(src)
            return pg.range_check(c, mn, mx, w) if gadget == "range_check" else pg.max_bound(c, mxs, w)[0]
        y = build(wit)                                    # device input: three launches over all instances
        assert c.check_circuit_satisfied() == (0, None)
        ref_y = y.values()
        vpi = 653 if gadget == "range_check" else 326
        picks = (0, n // 4 - 1, n // 4 + 1024, n // 2 + 5, n - 1)
        ref_vars = [c.variables(5 + n + vpi * i, vpi).copy() for i in picks]
        host = torch.zeros((n, 4), dtype=torch.int64).pin_memory()
        y = build(pinned)                                 # host input: chunk by chunk
        c.read_column_into(y, host, asynchronous=True)    # per chunk on the copy stream
        assert c.check_circuit_satisfied() == (0, None)
        c.sync()
        assert (host.numpy().view(np.uint64) == ref_y).all()
        assert (y.values() == ref_y).all()
        for i, ref in zip(picks, ref_vars):
            assert (c.variables(5 + n + vpi * i, vpi) == ref).all(), (gadget, i)
        # a partial asynchronous read takes the ordinary path
        part = torch.zeros((1000, 4), dtype=torch.int64).pin_memory()
        c.read_column_into(y, part, i0=n // 2, cnt=1000, asynchronous=True); c.sync()
        assert (part.numpy().view(np.uint64) == ref_y[n // 2: n // 2 + 1000]).all()


@pytest.mark.parametrize("mode", [pg.CHECK_GENERIC, pg.CHECK_SPARSE])
def test_check_modes_at_scale(oracle, torch_cuda, mode):
    """2^18 range_check instances: both evaluations give verdict 0, and both see two injected wrong claims."""
    torch = torch_cuda
    n = 1 << 18
    c = gpu_composer(check_mode=mode)
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit)
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    assert c.check_circuit_satisfied() == (0, None)
    claims = oracle.from_ints([1, 0, 1, 1, 0, 0] + [1, 0] * 5)   # instances 3 and 4 get the wrong claim
    sub = c.add_input(y.values(0, 16))                           # 16 result values re-allocated, then constrained
    c.constrain_to_constant(sub, claims)
    bad, first = c.check_circuit_satisfied()
    assert bad == 2 and first == 3 + 271 * n + 3


def test_empty_and_ragged_gpu(oracle):
    from tests.test_emu_engine import _empty_and_ragged
    _empty_and_ragged(gpu_composer, oracle)


# ---------------------------------------------------------------------------------------------------- evaluation domain (8f.2)
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 10, 11, 12, 14, 17, 20])
def test_fft_vs_oracle(oracle, log_n):
    """pg_fft (bit-reversal + shared-memory butterfly passes) against the restated serial FFT, forward and inverse; the sizes
    cover 1, 2 and 3 passes and every first-pass width."""
    c = gpu_composer()
    a = oracle.from_ints(synth_wide(70 + log_n, 1 << log_n)) if log_n <= 14 else None
    if a is None:                                               # larger inputs straight from the device generator
        import torch
        t = torch.empty((1 << log_n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 70 + log_n, 0, 0, t); c.sync()   # the engine runs on its own stream
        a = t.cpu().numpy().view(np.uint64)
    assert np.array_equal(c.fft(a), oracle.fft(a))
    assert np.array_equal(c.fft(a, inverse=True), oracle.fft(a, inverse=True))


def test_fft_golden_gpu(oracle, golden_domain):
    c = gpu_composer()
    for log_n, vec in golden_domain["fft"].items():
        a = oracle.from_ints([int(x, 16) for x in vec["input"]])
        assert [hx(v) for v in oracle.to_ints(c.fft(a))] == vec["fft"]
        assert [hx(v) for v in oracle.to_ints(c.fft(a, inverse=True))] == vec["ifft"]


def test_fft_properties_at_scale(oracle, torch_cuda):
    """2^24 scalars on the device (512 MiB per vector): round trip in place and out of place, the transform of a delta, and
    the decimation identity between two transform sizes."""
    torch = torch_cuda
    log_n = 24; n = 1 << log_n
    c = gpu_composer()
    x = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 77, 0, 0, x)
    fx = torch.empty_like(x)
    c.fft(x, out=fx); c.sync()                                  # the engine runs on its own stream: join before torch touches fx
    back = fx.clone(); torch.cuda.synchronize()
    c.fft(back, inverse=True); c.sync()                         # in place
    assert torch.equal(back, x)
    # decimation: the even-indexed outputs of a size-m FFT are the size-m/2 FFT of a_j + a_(j+m/2)
    m = 1 << 16
    s_np = x[:m].cpu().numpy().view(np.uint64)
    folded = c.fr_op(1, s_np[: m // 2], s_np[m // 2:])
    assert np.array_equal(c.fft(s_np)[0::2], c.fft(folded))
    delta = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    five = torch.from_numpy(oracle.from_ints([5]).view(np.int64)).cuda()
    delta[0] = five[0]; torch.cuda.synchronize()
    c.fft(delta); c.sync()
    assert bool((delta == five[0]).all())


def test_wire_polynomials_golden_gpu(oracle, golden, golden_domain):
    from oracle.gen_golden import coeff_digest
    hits = 0
    for name, spec in golden.items():
        if spec["expected"]["error"]:
            continue
        _s, oc = run_oracle(spec["program"], return_composer=True)
        _snap, c = run_engine(spec["program"], gpu_composer, oracle, return_composer=True)
        got = c.wire_polynomials()
        assert np.array_equal(got, oc.wire_polynomials()), name
        if name in golden_domain["wire_polynomials"]:
            assert coeff_digest([oracle.to_ints(got[w]) for w in range(4)]) == golden_domain["wire_polynomials"][name]["digest"]
            hits += 1
    assert hits >= 4


def test_wire_polynomials_batch_vs_oracle(oracle, torch_cuda):
    """1500 range_check instances (406 503 rows -> domain 2^19, two butterfly passes, tiled + ragged materialisation):
    coefficients equal to the oracle's ifft of the sequential composer's wire columns."""
    n = 1500
    wit = synth_wide(61, n)
    wit = [w % 2 ** 64 if i % 2 == 0 else w for i, w in enumerate(wit)]
    prog = [dict(op="add_input", values=[hx(w) for w in wit]), dict(op="range_check", min=hx(0), max=hx(2 ** 64), witness=0)]
    _s, oc = run_oracle(prog, return_composer=True)
    c = gpu_composer()
    w = c.add_input(oracle.from_ints(wit))
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    assert c.circuit_size() == oc.n and c.domain_log_size() == 19
    assert np.array_equal(c.wire_polynomials(), oc.wire_polynomials())


def test_wire_polynomials_at_scale(oracle, torch_cuda):
    """2^16 range_check instances: 17.8 M rows -> domain 2^25 (4 GiB of coefficients).  Evaluating the polynomials back over the
    domain (forward transform) must reproduce the materialised wire values on the rows and zeros on the padding."""
    torch = torch_cuda
    n = 1 << 16
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit)
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    rows = c.circuit_size(); k = c.domain_log_size(); N = 1 << k
    assert k == 25
    polys = torch.empty((4, N, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    c.wire_polynomials(out=polys)
    w_val = torch.empty((4, rows, 4), dtype=torch.int64, device="cuda")
    c._ok(c._L.pg_materialize_rows(c._ctx, 0, rows, None, C.c_void_p(w_val.data_ptr()), None, None, 1), "pg_materialize_rows")
    for col in range(4):
        c.fft(polys[col]); c.sync()
        assert torch.equal(polys[col, :rows], w_val[col]) and not bool(polys[col, rows:].any())


# ---------------------------------------------------------------------------------------------------- commitments (8f.2, second half)
def test_g1_group_law_gpu(oracle):
    from tests.test_emu_msm import to_engine, to_oracle
    c = gpu_composer()
    g = oracle.g1_generator()
    rnd = random.Random(8)
    ks = [1, 2, 3, Q - 1, Q - 2, rnd.randrange(Q), rnd.randrange(Q)]
    pts = oracle.g1_mul(np.repeat(g, len(ks), axis=0), oracle.from_ints(ks))
    inf = np.zeros((1, 13), dtype=np.uint64); inf[0, 12] = 1
    lhs = np.concatenate([pts[0:1], pts[1:2], pts[1:2], inf, pts[5:6], inf, pts[2:3]])
    rhs = np.concatenate([pts[1:2], pts[1:2], pts[4:5], pts[6:7], inf, inf, pts[3:4]])
    assert np.array_equal(to_oracle(c.g1_op(0, to_engine(lhs), to_engine(rhs))), oracle.g1_add(lhs, rhs))
    assert c.g1_op(1, to_engine(np.concatenate([pts, inf])))[:, 0].tolist() == [1] * (len(ks) + 1)
    sc = oracle.from_ints([0, 1, 5, Q - 1] + [rnd.randrange(Q) for _ in range(60)])
    assert np.array_equal(to_oracle(c.g1_fixed_base_mul(sc)), oracle.g1_mul(np.repeat(g, 64, axis=0), sc))
    beta = oracle.from_ints([rnd.randrange(Q)])
    assert np.array_equal(to_oracle(c.srs_powers(beta[0], 48)), oracle.srs_powers(beta, 48))
    # table of window multiples + batch normalisation (from 2^16 scalars on, or 256 once the generator's table exists), powers of beta
    # computed on the device
    big = oracle.from_ints([rnd.randrange(Q) for _ in range(1 << 16)])
    got = to_oracle(c.g1_fixed_base_mul(big))
    pick = [0, 1, 777, (1 << 16) - 1]
    assert np.array_equal(got[pick], oracle.g1_mul(np.repeat(g, len(pick), axis=0), big[pick]))
    n = 3000
    ks = [0, 1, 255, 256, 2 ** 248, Q - 1] + [rnd.randrange(Q) for _ in range(n - 6)]
    ks[77] = 0; ks[n - 1] = 0
    sc = oracle.from_ints(ks)
    assert np.array_equal(to_oracle(c.g1_fixed_base_mul(sc)), oracle.g1_mul(np.repeat(g, n, axis=0), sc))
    assert np.array_equal(to_oracle(c.srs_powers(beta[0], 2500)), oracle.srs_powers(beta, 2500))


@pytest.mark.parametrize("n", [1, 2, 7, 33, 100, 1000, 6000])
def test_msm_vs_oracle(oracle, n):
    """pg_msm (digits, CUB sort, bucket sums, chunked running sums, Horner) against the restated serial Pippenger."""
    from tests.test_emu_msm import to_engine, to_oracle
    c = gpu_composer()
    rnd = random.Random(n)
    srs = c.srs_powers(oracle.from_ints([rnd.randrange(Q)])[0], n)            # checked against the oracle in test_g1_group_law_gpu
    sc = synth_wide(40 + n, n)
    sc[0] = 1
    if n > 4:
        sc[2] = 0; sc[3] = 1; sc[4] = Q - 1
    if n > 40:
        srs[7] = srs[6]; sc[7] = sc[6]                                        # equal points in one bucket
        srs[9] = 0                                                            # a point at infinity
    got = c.msm(srs, oracle.from_ints(sc))
    assert np.array_equal(to_oracle(got.reshape(1, 12)), oracle.g1_msm(to_oracle(srs), oracle.from_ints(sc)))


def test_msm_edge_cases_gpu(oracle):
    from tests.test_emu_msm import to_engine, to_oracle
    c = gpu_composer()
    srs = c.srs_powers(oracle.from_ints([77])[0], 16)
    assert not c.msm(srs, oracle.from_ints([0] * 16)).any()
    assert not c.msm(np.zeros((0, 12), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64)).any()
    same = np.repeat(srs[3:4], 16, axis=0)
    assert np.array_equal(to_oracle(c.msm(same, oracle.from_ints([1] * 16)).reshape(1, 12)), oracle.g1_mul(to_oracle(srs[3:4]), oracle.from_ints([16])))
    assert not c.msm(np.concatenate([srs[3:4], srs[3:4]]), oracle.from_ints([5, Q - 5])).any()


@pytest.mark.parametrize("log_n", [18, 22])
def test_commitment_is_polynomial_at_beta(oracle, torch_cuda, log_n):
    """Size-independent check at 2^18 (15-bit windows) and 2^22 terms (16-bit windows), device-resident: against
    powers_of_g[i] = beta^i * G the commitment of a coefficient vector is poly(beta) * G -- the right-hand side needs one Horner
    evaluation on big ints and one fixed-base multiplication."""
    from tests.test_emu_msm import to_oracle
    torch = torch_cuda
    n = 1 << log_n
    c = gpu_composer()
    beta = 0x2b6cedcb87925c23c999e990f3f29c6d0748d9d99f59ff1105d314967254398f % Q
    srs = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    c.srs_powers(oracle.from_ints([beta])[0], n, out=srs)
    coeffs = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 91, 0, 0, coeffs); c.sync()
    got = c.msm(srs, coeffs)
    vals = oracle.to_ints(coeffs.cpu().numpy().view(np.uint64))
    acc = 0
    for v in reversed(vals):
        acc = (acc * beta + v) % Q
    want = c.g1_fixed_base_mul(oracle.from_ints([acc]))
    assert np.array_equal(got, want[0])
    assert np.array_equal(to_oracle(want), oracle.g1_mul(oracle.g1_generator(), oracle.from_ints([acc])))


def test_msm_split_consistency_at_scale(oracle, torch_cuda):
    """2^24 terms: the sum over all terms equals the sum of the two half-size sums, which in turn are anchored by the 2^22
    polynomial-at-beta check above."""
    torch = torch_cuda
    n = 1 << 24
    c = gpu_composer()
    pts = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    ks = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 92, 0, 0, ks); c.sync()
    # cheap distinct points: a 2^12-entry table of multiples of G, repeated (MSM does not care about repeated points)
    table = torch.from_numpy(c.srs_powers(oracle.from_ints([3])[0], 1 << 12).view(np.int64)).cuda()
    pts.view(n >> 12, 1 << 12, 12)[:] = table
    sc = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 93, 0, 0, sc); c.sync(); torch.cuda.synchronize()
    whole = c.msm(pts, sc)
    h = n // 2
    lo, hi = c.msm(pts[:h], sc[:h]), c.msm(pts[h:], sc[h:])
    assert np.array_equal(c.g1_op(0, lo.reshape(1, 12), hi.reshape(1, 12))[0], whole) and whole.any()


def test_commit_wire_polynomials_gpu(oracle, golden):
    """The four wire commitments of a 40-instance range_check batch (10 843 rows, domain 2^14) and of a golden program equal
    the oracle's msm_variable_base over the oracle's ifft of the sequential composer's wire columns."""
    from tests.test_emu_msm import to_oracle
    wit = synth_wide(62, 40)
    prog = [dict(op="add_input", values=[hx(w % 2 ** 64 if i % 2 == 0 else w) for i, w in enumerate(wit)]),
            dict(op="range_check", min=hx(0), max=hx(2 ** 64), witness=0)]
    for program in (golden["kat_max_bound_0_ok"]["program"], prog):
        _s, oc = run_oracle(program, return_composer=True)
        _snap, c = run_engine(program, gpu_composer, oracle, return_composer=True)
        k = c.domain_log_size()
        srs = c.srs_powers(oracle.from_ints([0x51ac582950405194])[0], 1 << k)
        got = c.commit_wire_polynomials(srs)
        polys = oc.wire_polynomials()
        for w in range(4):
            assert np.array_equal(to_oracle(got[w:w + 1]), oracle.g1_msm(to_oracle(srs), polys[w])), w


def test_g1_golden_gpu(oracle, golden, golden_g1):
    from tests.test_emu_msm import check_g1_golden
    check_g1_golden(gpu_composer, oracle, golden, golden_g1)


def test_fft_round_trip_2p27(oracle, torch_cuda):
    """2^27 scalars (4 GiB vector, 4 GiB scratch, 2 GiB twiddles; 4 passes): forward then inverse is the identity, and the even
    outputs equal FFT_(n/2)(low half) + FFT_(n/2)(high half) (decimation + linearity) on a strided sample of 4096 positions."""
    torch = torch_cuda
    n = 1 << 27
    c = gpu_composer()
    x = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 79, 0, 0, x); c.sync()
    y = x.clone(); torch.cuda.synchronize()
    c.fft(y); c.sync()
    idx = torch.arange(0, n // 2, (n // 2) >> 12, device="cuda")
    even = y[0::2][idx].cpu().numpy().view(np.uint64)
    c.fft(y, inverse=True); c.sync()
    assert torch.equal(y, x)
    del y
    lo, hi = x[: n // 2].clone(), x[n // 2:].clone(); torch.cuda.synchronize()
    c.fft(lo); c.fft(hi); c.sync()
    s = c.fr_op(1, lo[idx].cpu().numpy().view(np.uint64), hi[idx].cpu().numpy().view(np.uint64))
    assert np.array_equal(s, even)


def test_evaluation_form_commitments_gpu(oracle, golden, torch_cuda):
    """pg_commit_wire_evaluations (wire values x Lagrange-basis SRS, no FFT) returns the same four group elements as
    pg_commit_wire_polynomials (coefficients x monomial powers): on a golden program against the oracle too, and on a batch of
    1500 range_check instances (domain 2^19) with device-resident SRS forms."""
    from tests.test_emu_msm import to_oracle
    from oracle import pymodel as pm
    torch = torch_cuda
    beta_i = 0x1d0c3a5e7f9b2468ace013579bdf02468ace13579bdf048c159d26ae37bf48c1 % Q
    beta = oracle.from_ints([beta_i])
    c0 = gpu_composer()
    assert oracle.g1_to_ints(to_oracle(c0.srs_lagrange(beta[0], 4))) == pm.srs_lagrange(beta_i, 4)
    _s, oc = run_oracle(golden["kat_range_check_0_ok"]["program"], return_composer=True)
    _snap, c = run_engine(golden["kat_range_check_0_ok"]["program"], gpu_composer, oracle, return_composer=True)
    k = c.domain_log_size()
    mono = c.srs_powers(beta[0], 1 << k)
    got = c.commit_wire_evaluations(c.srs_lagrange(beta[0], k))
    assert np.array_equal(got, c.commit_wire_polynomials(mono))
    polys = oc.wire_polynomials()
    assert all(np.array_equal(to_oracle(got[w:w + 1]), oracle.g1_msm(to_oracle(mono), polys[w])) for w in range(4))
    # batch, device-resident
    n = 1500
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 2, 2, 64, wit)
    w = c.add_input(wit)
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    k = c.domain_log_size(); assert k == 19
    mono = torch.empty((1 << k, 12), dtype=torch.int64, device="cuda"); lag = torch.empty_like(mono); torch.cuda.synchronize()
    c.srs_powers(beta[0], 1 << k, out=mono); c.srs_lagrange(beta[0], k, out=lag)
    a, b = c.commit_wire_evaluations(lag), c.commit_wire_polynomials(mono)
    assert np.array_equal(a, b) and a.any()


def test_msm_skewed_scalars_gpu(oracle, torch_cuda):
    """2^22 terms whose scalars look like wire values (ones, zeros, tiny values, a few full-size ones): the digit-1 bucket of
    window 0 holds millions of entries.  Expected value by linearity over the 4096 distinct points."""
    from tests.test_emu_msm import to_oracle
    torch = torch_cuda
    n, m = 1 << 22, 1 << 12
    c = gpu_composer()
    table = c.srs_powers(oracle.from_ints([0xabcdef])[0], m)
    pts = torch.from_numpy(table.view(np.int64)).cuda().repeat(n // m, 1)
    i = np.arange(n, dtype=np.uint64)
    vals = np.ones(n, dtype=np.uint64)
    vals[i % 7 == 0] = 0
    small = i % 11 == 3
    vals[small] = (i[small] * np.uint64(2654435761)) % np.uint64(5)
    sc_int = vals.astype(object)
    big_idx = np.arange(5, n, 50001)
    bigs = synth_wide(78, len(big_idx))
    for k, j in enumerate(big_idx):
        sc_int[j] = bigs[k]
    # Montgomery form on the device: canonical limbs -> fr_op(to Montgomery) is host-sized, so build small/ones by table lookup
    lut = oracle.from_ints([0, 1, 2, 3, 4])
    sc = lut[np.minimum(vals, 4).astype(np.int64)].copy()
    sc[big_idx] = oracle.from_ints(bigs)
    sc_dev = torch.from_numpy(sc.view(np.int64)).cuda(); torch.cuda.synchronize()
    got = c.msm(pts, sc_dev)
    per_point = np.zeros(m, dtype=object)
    np.add.at(per_point, (i % m).astype(np.int64), sc_int)
    want = oracle.g1_msm(to_oracle(table), oracle.from_ints([int(v) % Q for v in per_point]))
    assert np.array_equal(to_oracle(got.reshape(1, 12)), want)


# ---- dusk-plonk's native range gate --------------------------------------------------------------------------
from tests import range_gate_cases as rgc  # noqa: E402


@pytest.mark.parametrize("bits", rgc.WIDTHS)
def test_range_gate_vs_oracle(oracle, bits):
    rgc.vs_oracle(gpu_composer, oracle, bits)


def test_range_gate_bad_arguments(oracle):
    rgc.bad_arguments(gpu_composer, oracle)


def test_range_gate_fault_injection(oracle):
    rgc.fault_injection(gpu_composer, oracle)


@pytest.mark.parametrize("mode", [pg.CHECK_GENERIC, pg.CHECK_SPARSE])
def test_range_gate_at_scale(oracle, torch_cuda, mode):
    """2^22 instances of the 64-bit range gate (10 rows, 32 accumulators each): even instances (uniform u64) fit, odd ones (uniform
    Fr) do not, so exactly the closing assert_equal rows of the odd instances fail; accumulators of sampled instances against the
    oracle; then the same on in-range witnesses only: satisfied."""
    torch = torch_cuda
    n = 1 << 22
    c = gpu_composer(check_mode=mode)
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    c.synth(SEED, 7, 2, 64, wit)
    w = c.add_input(wit)
    c.range_gate(w, 64)
    assert c.circuit_size() == 3 + 10 * n and c.num_variables() == 5 + n + 32 * n
    assert c.check_circuit_satisfied() == (n // 2, 3 + 10 + 9)
    wv = w.values()
    rng = random.Random(64)
    for i in [0, 1, n - 2, n - 1] + [rng.randrange(n) for _ in range(12)]:
        oc = oracle.Composer()
        oc.range_gate_batch(oc.add_input_batch(wv[i:i + 1]), 64)
        assert (oc.variables()[6:] == c.variables(5 + n + 32 * i, 32)).all(), f"accumulators of instance {i}"
        rows = c.rows(3 + 10 * i, 10, want=("w_idx",))
        o_w = oc.wires()[:, 3:]
        remap = np.where(o_w == 0, 0, np.where(o_w == 5, 5 + i, o_w - 6 + 5 + n + 32 * i)).astype(np.uint64)
        assert (rows["w_idx"] == remap).all(), f"wires of instance {i}"
    qa, qr = c.gate_selectors(3, 20)
    one = oracle.from_ints([1])[0]
    assert (qr[:8] == one).all() and (qr[8:10] == 0).all() and (qa[:9] == 0).all() and (qa[9] == one).all() and (qr[10:18] == one).all()
    c2 = gpu_composer(check_mode=mode)
    c2.synth(SEED, 8, 1, 64, wit)
    c2.range_gate(c2.add_input(wit), 64)
    assert c2.check_circuit_satisfied() == (0, None)


def test_range_gate_poked_witness(oracle):
    """Small segments: the row-parallel kernels."""
    rgc.poked_witness(gpu_composer, oracle)


def test_range_gate_poked_witness_large_segments(oracle):
    """Enough instances for the one-thread-per-instance kernels (k_check / k_check_prog) next to k_check_gates."""
    rgc.poked_witness(gpu_composer, oracle, n=148 * 320 + 77, trials=3, arith="maybe_equal")


# ---- the headline kernels must DETECT violations at their own launch shape --------------------------------
from tests import fault_cases as fc  # noqa: E402

HEADLINE_N = 148 * 320 + 77          # past run_check's 320 * SMs threshold: one thread per instance (k_check / k_check_prog)
KINDS = {pg.CHECK_GENERIC: "instance_generic", pg.CHECK_SPARSE: "program"}


def test_poked_range_check_headline_shape(oracle):
    """range_check k=65 (271 rows / 653 variables), 47 437 instances: V, A_j, U, Zv, Y of both decompositions and O poked at both
    ends, at warp / block boundaries and mid-batch; `k_check<GENERIC>` and the compiled row program `k_check_prog` must report exactly
    the rows the big-int evaluation of the dumped rows reports (count and first row).  Fails if either kernel is stubbed to return 0."""
    fc.poked_range_segment(gpu_composer, oracle, n=HEADLINE_N, gadget="range_check", bits=64, expect_kind=KINDS)


def test_poked_max_bound_headline_shape_per_instance_bounds(oracle):
    """max_bound with per-instance 64-bit bounds (q_c parameter slots): the same at the same launch shape."""
    fc.poked_range_segment(gpu_composer, oracle, n=HEADLINE_N, gadget="max_bound", bits=64, per_instance_bounds=True, expect_kind=KINDS, seed=6)


def test_poked_max_bound_k253(oracle):
    """C3's shape: max_bound, k = 253 (511 rows, 514 variables), large segment."""
    fc.poked_range_segment(gpu_composer, oracle, n=HEADLINE_N, gadget="max_bound", bits=252, expect_kind=KINDS, seed=7)


def test_fused_check_scalar_gadgets(oracle):
    """PG_F_FUSED_CHECK over maybe_equal / is_non_zero / conditionally_select_* at a size that takes the one-thread-per-instance kernels
    when a poke sends a segment back to them (tests/fault_cases.py)."""
    fc.fused_scalar_gadgets(gpu_composer, oracle, n=90)
    fc.fused_scalar_gadgets(gpu_composer, oracle, n=HEADLINE_N)


def test_fused_check_at_the_headline_launch_shape(oracle):
    """PG_F_FUSED_CHECK: the range gadgets' witness kernels evaluate the rows they generate (kind "fused": no check launch for the
    segment); overwriting a Variable sends the segment back to k_check_prog, which must report exactly the big-int verdict."""
    fc.poked_range_segment(gpu_composer, oracle, n=HEADLINE_N, gadget="range_check", bits=64, modes=(pg.CHECK_SPARSE,),
                           expect_kind={pg.CHECK_SPARSE: "program"}, fused=True)
    fc.poked_range_segment(gpu_composer, oracle, n=HEADLINE_N, gadget="max_bound", bits=64, per_instance_bounds=True, modes=(pg.CHECK_SPARSE,),
                           expect_kind={pg.CHECK_SPARSE: "program"}, fused=True, seed=8)


def test_is_non_zero_flags_gpu(oracle):
    fc.non_zero_flags_vs_oracle(gpu_composer, oracle, n=300)
    fc.non_zero_flags_vs_oracle(gpu_composer, oracle, n=1 << 16, zero_every=1024, mismatch_every=1024, modes=(pg.CHECK_GENERIC,))


def test_unreduced_inputs_rejected_gpu(oracle):
    fc.unreduced_inputs_rejected(gpu_composer, oracle)


def test_unreduced_input_in_chunked_host_copy(oracle, torch_cuda):
    """The ingest check of a chunked host copy runs chunk by chunk on the input stream: an unreduced scalar in the last chunk."""
    torch = torch_cuda
    n = (1 << 20) + 5
    c = gpu_composer()
    wit = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 2, 1, 64, wit); c.sync()
    host = wit.cpu().pin_memory()
    host[n - 3] = -1                                   # 2^256 - 1
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), c.add_input(host))
    with pytest.raises(pg.EngineError) as e:
        c.check_circuit_satisfied()
    assert e.value.code == -2 and f"index {n - 3}" in str(e.value)
    c.reset()
    host[n - 3] = 0
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), c.add_input(host))
    assert c.check_circuit_satisfied() == (0, None)


# ---- BASELINE.json configurations at their full sizes (size-independent properties) ----------------------------
def test_c3_full_size(oracle, torch_cuda):
    """C3: 2^22 max_bound instances, per-instance 252-bit bounds (k = 253): 2.14e9 rows, a 35 GB table."""
    torch = torch_cuda
    n = 1 << 22
    if torch.cuda.mem_get_info()[0] < 60 * 2 ** 30:
        pytest.skip("needs ~40 GB of free HBM")
    c = gpu_composer(check_mode=pg.CHECK_SPARSE)
    mx = torch.empty((n, 4), dtype=torch.int64, device="cuda"); wit = torch.empty_like(mx)
    c.synth(SEED, 31, 3, 252, mx); c.synth(SEED, 32, 2, 250, wit)
    y, k = pg.max_bound(c, mx, c.add_input(wit))
    assert k == 253 and c.circuit_size() == 3 + 511 * n and c.num_variables() == 5 + n + 514 * n
    assert c.check_circuit_satisfied() == (0, None)
    res = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.read_column_into(y, res); c.sync()
    one = torch.from_numpy(oracle.from_ints([1]).view(np.int64)).cuda()
    assert bool((res[0::2] == one).all())
    frac_in = float((res[1::2] == one).all(dim=1).float().mean())
    assert 0.26 < frac_in < 0.295                      # y = 1 iff (max-1-x mod q) fits 253 bits: probability 2^253 / q = 0.2764
    # one instance row by row against the oracle, and a fault in the last instance
    i = n - 1
    oc = oracle.Composer()
    wv = c.variables(5 + i, 1)
    bound = torch.empty((1, 4), dtype=torch.int64, device="cuda"); bound.copy_(mx[i:i + 1]); torch.cuda.synchronize()
    oc.max_bound_batch(bound.cpu().numpy().view(np.uint64), oc.add_input_batch(wv))
    assert (oc.variables()[6:] == c.variables(5 + n + 514 * i, 514)).all()
    var = 5 + n + 514 * i + 257 + 100                  # an accumulator in the middle of the chain
    old = c.variables(var, 1)[0].copy()
    c.poke_variable(var, oracle.from_ints([12345])[0])
    bad, first = c.check_circuit_satisfied()
    assert bad == 2 and first == 3 + 511 * i + 2 * 100 + 1      # the accumulate rows that write and read A_100
    c.close()


def test_c4_full_size_with_error_path(oracle, torch_cuda):
    """C4 as SURVEY.md 8d specifies it: 2^24 x (maybe_equal + is_non_zero), value uniform Fr with 1/1024 forced to zero (the
    NonExistingInverse path), value_assigned = value except 1/1024 mismatches, maybe_equal with 50 % a = b."""
    torch = torch_cuda
    n = 1 << 24
    c = gpu_composer()
    a = torch.empty((n, 4), dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
    c.synth(SEED, 41, 0, 0, a); c.synth(SEED, 42, 0, 0, b); c.sync()
    b[0::2] = a[0::2]
    a[5::1024] = 0                                     # forced zeros (b differs there unless even: index 5 mod 1024 is odd)
    assigned = a.clone()
    assigned[600::1024, 0] ^= 1                        # mismatching value_assigned
    flags = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    va, vb = c.add_input(a), c.add_input(b)
    eq = pg.maybe_equal(c, va, vb)
    pg.is_non_zero_flags(c, va, assigned, layout=pg.NZ_UNIFORM, flags_out=flags)
    assert c.last_n_err == n // 1024
    bad, first = c.check_circuit_satisfied()
    c.sync()
    assert int(flags.sum()) == n // 1024 and bool(flags[5::1024].all())
    assert bad == 3 * (n // 1024) and first == 3 + 3 * n + 3 * 5 + 2      # zero: var*inv - 1; mismatch: assert_equal and var*inv - 1
    r = torch.empty((n, 4), dtype=torch.int64, device="cuda"); c.read_column_into(eq, r); c.sync()
    one = torch.from_numpy(oracle.from_ints([1]).view(np.int64)).cuda()
    assert bool((r[0::2] == one).all()) and bool((r[1::2] == 0).all())
    # the `?` variant stops at the first zero
    c.reset()
    va = c.add_input(a)
    with pytest.raises(pg.NonExistingInverse) as e:
        pg.is_non_zero(c, va, a)
    assert (e.value.n_err, e.value.first_err) == (n // 1024, 5) and c.circuit_size() == 3 + 3 * 5 + 1
    c.close()


def test_c5_full_size(oracle, torch_cuda):
    """C5: the mixed circuit of 2^26 rows (range_check k=65 / max_bound k=253 / is_non_zero / select_one + select_zero, a quarter
    of the rows each) on one GPU, both check modes: counts, verdict, result patterns."""
    torch = torch_cuda
    q = 1 << 24
    n_rc, n_mb, n_nz, n_sel = q // 271, q // 511, q // 3, q // 5
    one = torch.from_numpy(oracle.from_ints([1]).view(np.int64)).cuda()
    for mode in (pg.CHECK_GENERIC, pg.CHECK_SPARSE):
        c = gpu_composer(check_mode=mode)
        x_rc = torch.empty((n_rc, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 51, 2, 64, x_rc)
        x_mb = torch.empty((n_mb, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 52, 2, 250, x_mb)
        x_nz = torch.empty((n_nz, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 53, 0, 0, x_nz)
        x_sel = torch.empty((n_sel, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 54, 0, 0, x_sel)
        s_sel = torch.empty((n_sel, 4), dtype=torch.int64, device="cuda"); c.synth(SEED, 55, 1, 1, s_sel)
        y_rc = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), c.add_input(x_rc))
        y_mb, k = pg.max_bound(c, oracle.from_ints([2 ** 252]), c.add_input(x_mb))
        pg.is_non_zero(c, c.add_input(x_nz), x_nz)
        x, s = c.add_input(x_sel), c.add_input(s_sel)
        y1 = pg.conditionally_select_one(c, x, s)
        y0 = pg.conditionally_select_zero(c, y1, s)
        rows = 3 + 271 * n_rc + 511 * n_mb + 3 * n_nz + 5 * n_sel
        assert k == 253 and c.circuit_size() == rows and abs(rows - (1 << 26)) < 2000
        assert c.check_circuit_satisfied() == (0, None)
        r = y_rc.values()
        assert (r[0::2] == oracle.from_ints([1])[0]).all() and (r[1::2] == 0).all()
        # select_one(x, s) = s ? x : 1 ; select_zero(y, s) = y * s
        got1 = torch.empty((n_sel, 4), dtype=torch.int64, device="cuda"); c.read_column_into(y1, got1)
        got0 = torch.empty((n_sel, 4), dtype=torch.int64, device="cuda"); c.read_column_into(y0, got0); c.sync()
        is_one = (s_sel == one).all(dim=1)
        assert bool((got1[is_one] == x_sel[is_one]).all()) and bool((got1[~is_one] == one).all())
        assert bool((got0[is_one] == x_sel[is_one]).all()) and bool((got0[~is_one] == 0).all())
        c.close()
