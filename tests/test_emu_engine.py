"""CPU tests of the engine's HOST LOGIC (templates, Variable/row numbering, operand resolution, witness arithmetic)
through tests/emu: the same engine.hpp/bodies.cuh compiled by g++ over a loop backend.  Test infrastructure only -- the
CUDA kernels themselves are tested on the B200 by the `-m gpu` tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import plonk_gadgets_b200 as pg
from plonk_gadgets_b200 import _lib
from tests.engine_runner import run_engine
from tests.programs import Q, hx, run_oracle, synth_wide

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
ROOT = os.path.dirname(HERE)


def _build(name, src):
    out = os.path.join(EMU_DIR, "_build", name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [os.path.join(EMU_DIR, src)] + [os.path.join(ROOT, "plonk_gadgets_b200", "csrc", f)
                                          for f in os.listdir(os.path.join(ROOT, "plonk_gadgets_b200", "csrc"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(EMU_DIR, src)], check=True)
    return out


@pytest.fixture(scope="module")
def emu():
    return _lib.bind(C.CDLL(_build("libpg_emu.so", "engine_emu.cpp")))


@pytest.fixture(scope="module")
def fr_emu():
    return C.CDLL(_build("libfr_emu.so", "fr_emu.cpp"))


def test_emu_exports_the_abi(emu):
    assert emu.pg_abi_version() == _lib.ABI_VERSION


def test_emu_matches_golden(emu, golden, oracle):
    for name, spec in golden.items():
        exp = spec["expected"]
        snap = run_engine(spec["program"], lambda: pg.StandardComposer(_cdll=emu), oracle)
        assert (snap.n_rows, snap.n_vars) == (exp["n_rows"], exp["n_vars"]), name
        assert snap.unsat == exp["unsat"], name
        assert (list(snap.error) if snap.error else None) == exp["error"], name
        assert snap.digest() == exp["digest"], name
        for k, vals in exp["results"].items():
            assert [hx(v) for v in snap.results(int(k))] == vals, (name, k)


@pytest.mark.parametrize("bits", [1, 2, 31, 32, 33, 64, 65, 127, 128, 200, 252, 253])
def test_emu_range_check_vs_oracle(emu, oracle, bits):
    r = synth_wide(50 + bits, 12)
    mx = ((r[0] % 2 ** (bits - 1)) | 2 ** (bits - 1)) + 1 if bits > 1 else 2
    mn = r[1] % mx
    wit = [mn, mx - 1, mx, (mn - 1) % Q, 0, Q - 1, r[2], r[3] % mx, (r[4] % mx + mn) % Q, 2 ** bits, 2 ** bits - 1, 1]
    prog = [dict(op="add_input", values=[hx(x) for x in wit]), dict(op="range_check", min=hx(mn), max=hx(mx), witness=0),
            dict(op="max_bound", max=hx(mx), witness=0)]
    so = run_oracle(prog)
    se = run_engine(prog, lambda: pg.StandardComposer(_cdll=emu), oracle)
    assert se.digest() == so.digest()
    assert se.unsat == so.unsat == []


def test_emu_mixed_bits_rejected(emu, oracle):
    c = pg.StandardComposer(_cdll=emu)
    w = c.add_input(oracle.from_ints([1, 2]))
    with pytest.raises(pg.EngineError) as e:
        pg.range_check(c, oracle.from_ints([0, 0]), oracle.from_ints([2 ** 10, 2 ** 20]), w)
    assert e.value.code == -4
    assert c.circuit_size() == 3 and c.num_variables() == 5 + 2     # nothing was appended
    assert c.check_circuit_satisfied() == (0, None)


def test_emu_fault_injection(emu, oracle):
    """Flip one materialised wire value: exactly the rows that read it become unsatisfied (L4 of SURVEY.md 4.5)."""
    c = pg.StandardComposer(_cdll=emu)
    w = c.add_input(oracle.from_ints([12345, 2 ** 70]))
    pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), w)
    rows = c.rows()
    assert c.check_rows(rows["w_val"], rows["sel"], rows["pi"]) == (0, None)
    bad_w = rows["w_val"].copy()
    bad_w[2, 100] = oracle.from_ints([424242])[0]      # output wire of row 100
    assert c.check_rows(bad_w, rows["sel"], rows["pi"]) == (1, 100)


def test_fr_even_odd_multiplier_host_emulation(fr_emu):
    """The device multiplier's even/odd carry-chain algorithm (fr.cuh) through its host emulation, against big ints; and no
    dropped carry is ever non-zero."""
    import random
    R = (1 << 256) % Q
    rinv = pow(R, -1, Q)
    rng = random.Random(5)
    edge = [0, 1, 2, Q - 1, Q - 2, R, Q - R, 2 ** 32 - 1, 2 ** 64 - 1, (1 << 254) + 12345, Q >> 1, (0xffffffff << 224) % Q, Q - 2 ** 224]
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(Q), rng.randrange(Q)) for _ in range(3000)]
    pack = lambda vs: np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vs), dtype=np.uint64).reshape(-1, 4).copy()
    unpack = lambda a: [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
    A, B = pack([p[0] for p in pairs]), pack([p[1] for p in pairs])
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for fn in ("emu_mul_eo", "emu_mul_cios"):
        out = np.zeros_like(A)
        getattr(fr_emu, fn)(C.c_uint64(len(pairs)), vp(A), vp(B), vp(out))
        assert unpack(out) == [a * b * rinv % Q for a, b in pairs], fn
    assert fr_emu.emu_violations() == 0
    # the SCANNED (second) operand of the multiplier may be ANY 256-bit value when the first is below q (the running value stays below
    # 2q): the range widget's rows pass f - 3 as the unreduced sum f + (q - 3) (bodies.cuh range_row_holds)
    wide = [2 ** 256 - 1, 2 ** 256 - 2 ** 32, 2 * Q - 1, 2 * Q - 2, Q, Q + 1] + [rng.randrange(2 ** 256) for _ in range(2000)]
    small = [Q - 1, Q - 2, 1, 0, R, Q >> 1] + [rng.randrange(Q) for _ in range(2000)]
    W, S = pack(wide), pack(small)
    out = np.zeros_like(S)
    fr_emu.emu_mul_eo(C.c_uint64(len(small)), vp(S), vp(W), vp(out))
    assert unpack(out) == [a * b * rinv % Q for a, b in zip(small, wide)]
    assert fr_emu.emu_violations() == 0
    for fn, f in (("emu_add", lambda a, b: (a + b) % Q), ("emu_sub", lambda a, b: (a - b) % Q)):
        out = np.zeros_like(A)
        getattr(fr_emu, fn)(C.c_uint64(len(pairs)), vp(A), vp(B), vp(out))
        assert unpack(out) == [f(a, b) for a, b in pairs], fn


def test_fr_inversions_host_emulation(fr_emu):
    """fr_inv_binary (binary extended Euclid, the batch inversion's one inversion per block) and fr_inv_fermat (x^(q-2)) return
    the same limbs, and both are the big-int inverse: Montgomery form in, Montgomery form out."""
    import random
    R = (1 << 256) % Q
    rng = random.Random(11)
    vals = [1, 2, 3, 4, Q - 1, Q - 2, R, Q - R, (Q + 1) // 2, 2 ** 32, 2 ** 32 - 1, 2 ** 64 - 1, 2 ** 254, 2 ** 254 + 1, 2 ** 255 % Q,
            pow(R, -1, Q), pow(2, -1, Q) * R % Q] + [2 ** i for i in range(0, 254, 17)] + [rng.randrange(1, Q) for _ in range(2000)]
    pack = lambda vs: np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vs), dtype=np.uint64).reshape(-1, 4).copy()
    unpack = lambda a: [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
    A = pack(vals)                                     # stored limbs x~ = x R  =>  x = x~ / R
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    expect = [pow(v * pow(R, -1, Q) % Q, -1, Q) * R % Q for v in vals]
    for fn in ("emu_inv_binary", "emu_inv"):
        out = np.zeros_like(A)
        getattr(fr_emu, fn)(C.c_uint64(len(vals)), vp(A), vp(out))
        assert unpack(out) == expect, fn


def test_fp_inversions_host_emulation(fr_emu):
    """fp_inv (binary extended Euclid; the end of every MSM, g1x_to_affine) against x^(p-2) and against big ints, Montgomery
    form (R = 2^384) in and out."""
    import random
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    R = (1 << 384) % P
    rng = random.Random(13)
    vals = [1, 2, 3, P - 1, P - 2, R, P - R, (P + 1) // 2, 2 ** 32, 2 ** 64 - 1, 2 ** 380, 2 ** 380 + 1, pow(R, -1, P)] \
        + [2 ** i for i in range(0, 381, 29)] + [rng.randrange(1, P) for _ in range(600)]
    pack = lambda vs: np.frombuffer(b"".join(v.to_bytes(48, "little") for v in vs), dtype=np.uint32).reshape(-1, 12).copy()
    unpack = lambda a: [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
    A = pack(vals)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    expect = [pow(v * pow(R, -1, P) % P, -1, P) * R % P for v in vals]
    for fn in ("emu_fp_inv", "emu_fp_inv_fermat"):
        out = np.zeros_like(A)
        getattr(fr_emu, fn)(C.c_uint64(len(vals)), vp(A), vp(out))
        assert unpack(out) == expect, fn


def test_emu_wire_format(emu, oracle):
    """to_bytes / from_bytes (canonical LE 32 bytes) against the oracle's, including rejection of encodings >= q."""
    import random
    rng = random.Random(3)
    vals = [0, 1, Q - 1, 2 ** 255 % Q] + [rng.randrange(Q) for _ in range(50)]
    c = pg.StandardComposer(_cdll=emu)
    m = oracle.from_ints(vals)
    raw = c.to_bytes(m)
    assert [int.from_bytes(raw[i].tobytes(), "little") for i in range(len(vals))] == vals
    back, bad, first = c.from_bytes(raw)
    assert bad == 0 and first is None and (back == m).all()
    bogus = raw.copy()
    bogus[7] = np.frombuffer(Q.to_bytes(32, "little"), dtype=np.uint8)            # == q: rejected
    bogus[9] = 0xFF                                                                 # 2^256-1: rejected
    back, bad, first = c.from_bytes(bogus)
    assert (bad, first) == (2, 7) and (back[7] == 0).all() and (back[9] == 0).all() and (back[8] == m[8]).all()


def test_emu_permutation_map(emu, golden, oracle):
    """pg_permutation (copy-constraint cycles) against the oracle's perm.variable_map on every golden program, whole range and
    a sub-range."""
    from tests.programs import expected_sigma
    for name, spec in golden.items():
        so, oc = run_oracle(spec["program"], return_composer=True)
        se, c = run_engine(spec["program"], lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
        exp = expected_sigma(oc)
        got = c.permutation()
        assert got.shape == exp.shape and (got == exp).all(), name
        if so.n_rows > 10:
            assert (c.permutation(5, so.n_rows - 9) == exp[:, 5:so.n_rows - 4]).all(), name


def test_emu_permutation_same_column_twice(emu, oracle):
    """maybe_equal(a, a) / select(x, x): both operands are one Variable, its uses must form one cycle."""
    from tests.programs import expected_sigma
    prog = [dict(op="add_input", values=[hx(5), hx(0), hx(1)]), dict(op="maybe_equal", a=0, b=0), dict(op="select_one", y=1, select=1),
            dict(op="select_zero", x=0, select=0), dict(op="constrain_to_constant", a=1, constant=hx(1))]
    so, oc = run_oracle(prog, return_composer=True)
    se, c = run_engine(prog, lambda: pg.StandardComposer(_cdll=emu), oracle, return_composer=True)
    assert se.digest() == so.digest()
    assert (c.permutation() == expected_sigma(oc)).all()


def _materialize_vs_oracle(make_composer, oracle, n=70):
    """Whole-instance ranges go through the tiled path (>= 32 instances), ragged ends through the simple body: every column of
    pg_materialize_rows must equal the oracle's composer columns for arbitrary sub-ranges."""
    vals = synth_wide(21, n)
    wit = [v % 300 for v in vals]
    c = make_composer()
    w = c.add_input(oracle.from_ints(wit))
    mx = oracle.from_ints([200 + (i % 50) for i in range(n)])            # per-instance bounds, all 8 bits wide -> q_c parameters
    mn = oracle.from_ints([i % 7 for i in range(n)])
    y = pg.range_check(c, mn, mx, w)
    c.constrain_to_constant(y, oracle.from_ints([1]), oracle.from_ints(list(range(n))))   # per-instance PI
    oc = oracle.Composer()
    ow = oc.add_input_batch(oracle.from_ints(wit))
    oy = oc.range_check_batch(mn, mx, ow)
    oc.constrain_to_constant_batch(oy, oracle.from_ints([1]), oracle.from_ints(list(range(n))))
    assert c.circuit_size() == oc.n
    o_w, o_sel, o_pi, o_vars = oc.wires(), oc.selectors()[:6], oc.dense_pi(), oc.variables()
    total = oc.n
    for row0, cnt in ((0, total), (3, total - 3), (10, total - 25), (3 + 47 * 5 + 11, 47 * 40 + 5), (total - 60, 60)):
        rows = c.rows(row0, cnt)
        sl = slice(row0, row0 + cnt)
        assert (rows["w_idx"] == o_w[:, sl]).all(), (row0, cnt)
        assert (rows["sel"] == o_sel[:, sl]).all(), (row0, cnt)
        assert (rows["pi"] == o_pi[sl]).all(), (row0, cnt)
        assert (rows["w_val"] == o_vars[o_w[:, sl].astype(np.int64)]).all(), (row0, cnt)


def test_emu_materialize_ranges(emu, oracle):
    _materialize_vs_oracle(lambda: pg.StandardComposer(_cdll=emu), oracle)


@pytest.mark.parametrize("mode", [pg.CHECK_GENERIC, pg.CHECK_SPARSE])
def test_emu_check_modes_agree(emu, golden, oracle, mode):
    """Both evaluations of the gate equation give the oracle's verdict (satisfied and unsatisfied programs)."""
    for name in ("batch_mixed_circuit", "batch_max_bound_k8_claims", "kat_range_check_1_wrongclaim", "kat_is_non_zero_mismatch",
                 "batch_range_check_k65_per_instance_bounds", "kat_select_one_sel1", "batch_is_non_zero_maybe_equal"):
        spec = golden[name]
        snap = run_engine(spec["program"], lambda: pg.StandardComposer(check_mode=mode, _cdll=emu), oracle)
        assert snap.unsat == spec["expected"]["unsat"], name


def test_emu_degenerate_num_bits(emu, oracle):
    """bitlen(max-1) = 255: 2^255 wraps mod q and the reference's num_bits_closest_power_of_two returns bitlen(2^255 mod q)
    (SURVEY.md 8a row a7).  The engine must follow the reference into that corner, uniform and per-instance bounds alike."""
    mx = 2 ** 254 + 12345
    k_ref = (2 ** 255 % Q).bit_length()
    prog = [dict(op="add_input", values=[hx(5), hx(2 ** 254), hx(Q - 1)]), dict(op="max_bound", max=hx(mx), witness=0),
            dict(op="range_check", min=[hx(1), hx(2), hx(3)], max=[hx(mx), hx(mx + 1), hx(Q - 1)], witness=0)]
    so = run_oracle(prog)
    se = run_engine(prog, lambda: pg.StandardComposer(_cdll=emu), oracle)
    assert se.digest() == so.digest() and se.unsat == so.unsat == []
    assert so.n_rows == 3 + (2 * k_ref + 5) * 3 + (4 * k_ref + 11) * 3


def _empty_and_ragged(make_composer, oracle):
    """Empty batches append nothing; ragged batch sizes (1, 31, 33, 257) hit every tail path of the kernels."""
    c = make_composer()
    e = c.add_input(np.empty((0, 4), dtype=np.uint64))
    y = pg.range_check(c, oracle.from_ints([0]), oracle.from_ints([2 ** 64]), e)
    m = pg.maybe_equal(c, e, y)
    pg.is_non_zero(c, e, np.empty((0, 4), dtype=np.uint64))
    c.constrain_to_constant(m, oracle.from_ints([1]))
    assert (y.n, m.n) == (0, 0) and c.circuit_size() == 3 and c.num_variables() == 5
    assert c.check_circuit_satisfied() == (0, None)
    assert c.rows()["w_idx"].shape == (4, 3) and y.values().shape == (0, 4)
    for n in (1, 31, 33, 257):
        vals = [v % 2 ** 70 for v in synth_wide(n, n)]
        prog = [dict(op="add_input", values=[hx(v) for v in vals]), dict(op="range_check", min=hx(3), max=hx(2 ** 64), witness=0),
                dict(op="maybe_equal", a=0, b=1), dict(op="select_one", y=0, select=1)]
        so = run_oracle(prog)
        se = run_engine(prog, make_composer, oracle)
        assert se.digest() == so.digest() and se.unsat == so.unsat == [], n


def test_emu_empty_and_ragged(emu, oracle):
    _empty_and_ragged(lambda: pg.StandardComposer(_cdll=emu), oracle)


def test_emu_sparse_program_on_batches(emu, oracle):
    """Batches large enough for the one-thread-per-instance mapping, i.e. for the compiled structure-aware row program
    (SparseProgBody): the verdict (count and first bad row) must equal the generic evaluation's, on satisfied circuits and with
    wrong claims, wrong witnesses for is_non_zero, per-instance bounds (q_c parameters) and public inputs."""
    n = 64
    wit = synth_wide(55, n)
    wit = [w % 2 ** 40 if i % 2 == 0 else w for i, w in enumerate(wit)]
    mx = [(2 ** 39 | (w % 2 ** 39)) + 1 for w in synth_wide(56, n)]
    mn = [m // 3 for m in mx]
    claims = [1 if mn[i] <= wit[i] < mx[i] else 0 for i in range(n)]
    claims[5] ^= 1; claims[40] ^= 1                                             # two wrong claims
    prog = [dict(op="add_input", values=[hx(w) for w in wit]),
            dict(op="range_check", min=[hx(v) for v in mn], max=[hx(v) for v in mx], witness=0),
            dict(op="constrain_to_constant", a=1, constant=[hx(c) for c in claims]),
            dict(op="maybe_equal", a=0, b=1),
            dict(op="select_one", y=0, select=1),
            dict(op="select_zero", x=0, select=1),
            dict(op="constrain_to_constant", a=5, constant=hx(0), pi=[hx(-(0 if c else w)) for c, w in zip(claims, wit)]),
            dict(op="is_non_zero", var=0, assigned=[hx(w if i != 9 else w + 1) for i, w in enumerate(wit)])]
    ref = run_oracle(prog)
    assert len(ref.unsat) >= 3
    for mode in (pg.CHECK_GENERIC, pg.CHECK_SPARSE):
        snap = run_engine(prog, lambda: pg.StandardComposer(check_mode=mode, _cdll=emu), oracle)
        assert snap.unsat == ref.unsat and snap.digest() == ref.digest(), mode


# ---- dusk-plonk's native range gate --------------------------------------------------------------------------
from tests import range_gate_cases as rgc  # noqa: E402


@pytest.mark.parametrize("bits", rgc.WIDTHS)
def test_emu_range_gate_vs_oracle(emu, oracle, bits):
    rgc.vs_oracle(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle, bits)


def test_emu_range_gate_bad_arguments(emu, oracle):
    rgc.bad_arguments(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle)


def test_emu_range_gate_fault_injection(emu, oracle):
    rgc.fault_injection(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle)


def test_emu_range_gate_poked_witness(emu, oracle):
    rgc.poked_witness(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle)
    # 60 instances: the per-instance walk of the rows (GateRowsCheckBody::run: the next row's fourth wire carried as d_next / d)
    rgc.poked_witness(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle, n=60, trials=8, seed=12)


# ---- faults inside the range gadgets' own segments, per-instance Results, ingest checks (shared with the GPU tests) -----------
from tests import fault_cases as fc  # noqa: E402


@pytest.mark.parametrize("gadget,bits,per_inst", [("range_check", 8, False), ("max_bound", 12, True), ("range_check", 64, False)])
def test_emu_poked_range_segment(emu, oracle, gadget, bits, per_inst):
    """Host emulation of the per-instance check bodies (CheckBody::run<0>, SparseProgBody over the compiled row program): a
    poked Variable of a range segment makes exactly the rows a big-int evaluation names unsatisfied."""
    fc.poked_range_segment(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle, n=150 if bits < 64 else 131, gadget=gadget, bits=bits,
                           per_instance_bounds=per_inst, expect_kind={pg.CHECK_GENERIC: "instance_generic", pg.CHECK_SPARSE: "program"})


def test_emu_fused_check(emu, golden, oracle):
    """PG_F_FUSED_CHECK: rows evaluated inside witness generation give the verdicts of the separate check, and a poked segment
    falls back to the check kernel."""
    for name in ("batch_mixed_circuit", "batch_max_bound_k8_claims", "kat_range_check_1_wrongclaim", "batch_range_check_k65_per_instance_bounds",
                 "batch_is_non_zero_maybe_equal", "kat_is_non_zero_mismatch", "kat_select_one_sel1", "kat_select_zero_sel1_claim0",
                 "kat_maybe_equal_20_3330_wrongclaim", "batch_range_gate_mixed"):
        spec = golden[name]
        snap = run_engine(spec["program"], lambda: pg.StandardComposer(check_mode=pg.CHECK_SPARSE, fused_check=True, _cdll=emu), oracle)
        assert snap.unsat == spec["expected"]["unsat"], name
        assert snap.digest() == spec["expected"]["digest"], name
    fc.fused_scalar_gadgets(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle, n=90)
    for gadget, bits, per_inst in (("range_check", 8, False), ("max_bound", 12, True), ("range_check", 64, False)):
        fc.poked_range_segment(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle, n=131, gadget=gadget, bits=bits, per_instance_bounds=per_inst,
                               modes=(pg.CHECK_SPARSE,), expect_kind={pg.CHECK_SPARSE: "program"}, fused=True)


@pytest.mark.parametrize("mode", [pg.CHECK_GENERIC, pg.CHECK_SPARSE])
def test_emu_template_cache_reuse(emu, golden, oracle, mode):
    """A ctx keeps built range templates and resolved segment images across composer resets (engine.hpp range_cache): programs
    repeated and interleaved on ONE ctx must give the states fresh contexts give -- same bounds on another operand column, other
    bounds on the same tables, per-instance bounds, and a poke in between (the cache holds structure, never values)."""
    shared = pg.StandardComposer(check_mode=mode, _cdll=emu)
    def reuse():
        shared.reset()
        return shared
    names = ["batch_range_check_k65", "batch_max_bound_k8_claims", "batch_range_check_k65", "batch_range_check_k65_per_instance_bounds",
             "batch_mixed_circuit", "kat_range_check_1_wrongclaim", "batch_max_bound_k8_claims", "batch_mixed_circuit", "kat_range_check_0_ok",
             "kat_range_check_1_ok", "batch_range_check_k65"]
    for name in names:
        spec = golden[name]
        snap = run_engine(spec["program"], reuse, oracle)
        assert snap.unsat == spec["expected"]["unsat"], name
        assert snap.digest() == spec["expected"]["digest"], name
    # the same bounds over two different operand columns of one segment layout
    xs = [3, 200, 70000, 2 ** 20]
    for first in (True, False, True):
        shared.reset()
        a = shared.add_input(oracle.from_ints(xs)); b = shared.add_input(oracle.from_ints([x + 1 for x in xs]))
        y = pg.range_check(shared, oracle.from_ints([100]), oracle.from_ints([70001]), a if first else b)
        want = [1 if 100 <= (x if first else x + 1) < 70001 else 0 for x in xs]
        assert oracle.to_ints(y.values()) == want
        assert shared.check_circuit_satisfied() == (0, None)
        w_idx = shared.rows(3, 4 * 18 + 11, want=("w_idx",))["w_idx"]
        assert int(w_idx[0, 0]) == (5 if first else 5 + len(xs))          # the first row's wire is the operand's Variable


def test_emu_is_non_zero_flags(emu, oracle):
    fc.non_zero_flags_vs_oracle(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle)


def test_emu_unreduced_inputs_rejected(emu, oracle):
    fc.unreduced_inputs_rejected(lambda **kw: pg.StandardComposer(_cdll=emu, **kw), oracle)
