"""Generates tests/golden/programs.json from the independent big-int model (oracle/pymodel.py).

Run from the repo root:   python -m oracle.gen_golden
TEST INFRASTRUCTURE ONLY.  The golden file pins: row/variable counts, the list of unsatisfied rows (the verdict),
the values of every variable a call returns, and a SHA-256 digest of the complete composer state (all variables, all
four wire columns, all eleven selector columns, dense public inputs) in canonical little-endian encoding.

The first block of programs replays the reference's own known-answer tests (verdict-level pins):
  /root/reference/tests/range_gadgets_tests.rs:57-78 (max_bound), :120-169 (range_check),
  /root/reference/tests/scalar_gadgets_tests.rs:36,:53 (maybe_equal), :106,:119 (select_zero), :145-177 (select_one),
  :199,:205-235 (is_non_zero), /root/reference/src/range.rs:214-218 (decomposition of -100 in 8 bits).
The reference draws its "random" scalars from thread_rng; here they come from the seeded generator in tests/programs.py.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tests.programs import Q, hx, run_pymodel, synth_wide  # noqa: E402


def kat_programs():
    progs = {}
    # ---- max_bound KATs (tests/range_gadgets_tests.rs:57-78): claim = expected (must be satisfied) and negated claim
    mb = [(2 ** 128 - 1, 2 ** 127, True), (200, 100, True), (100, 200, False), (2 ** 128 - 1, 2 ** 130, False)]
    for i, (mx, w, exp) in enumerate(mb):
        for claim in (exp, not exp):
            progs[f"kat_max_bound_{i}_{'ok' if claim == exp else 'wrongclaim'}"] = dict(
                satisfied=(claim == exp),
                program=[dict(op="add_input", values=[hx(w)]), dict(op="max_bound", max=hx(mx), witness=0),
                         dict(op="constrain_to_constant", a=1, constant=hx(int(claim)))])
    # ---- range_check KATs (tests/range_gadgets_tests.rs:120-169)
    rc = [(50_000, 250_000, 50_001, True), (50_000, 250_000, 250_001, False), (50_000, 250_000, 250_000, False),
          (50_000, 250_000, 249_000, True), (50_000, 250_000, 50_000, True), (50_000, 250_000, 49_999, False),
          (2 ** 126, 2 ** 127 + 1, 2 ** 127 - 1, True), (50_000, 250_000, 18_598, False)]
    for i, (mn, mx, w, exp) in enumerate(rc):
        for claim in (exp, not exp):
            progs[f"kat_range_check_{i}_{'ok' if claim == exp else 'wrongclaim'}"] = dict(
                satisfied=(claim == exp),
                program=[dict(op="add_input", values=[hx(w)]), dict(op="range_check", min=hx(mn), max=hx(mx), witness=0),
                         dict(op="constrain_to_constant", a=1, constant=hx(int(claim)))])
    # ---- maybe_equal (tests/scalar_gadgets_tests.rs:36,:53)
    for a, b, exp in ((100, 100, True), (20, 3330, False)):
        for claim in (exp, not exp):
            progs[f"kat_maybe_equal_{a}_{b}_{'ok' if claim == exp else 'wrongclaim'}"] = dict(
                satisfied=(claim == exp),
                program=[dict(op="add_input", values=[hx(a)]), dict(op="add_input", values=[hx(b)]),
                         dict(op="maybe_equal", a=0, b=1), dict(op="constrain_to_constant", a=2, constant=hx(int(claim)))])
    r = synth_wide(100, 8)
    # ---- conditionally_select_zero (:106 selector 0 -> 0 ok ; :119 selector 1, random value, claim "0" rejected)
    progs["kat_select_zero_sel0"] = dict(satisfied=True, program=[
        dict(op="add_input", values=[hx(r[0])]), dict(op="add_input", values=[hx(0)]),
        dict(op="select_zero", x=0, select=1), dict(op="constrain_to_constant", a=2, constant=hx(0))])
    progs["kat_select_zero_sel1_claim0"] = dict(satisfied=False, program=[
        dict(op="add_input", values=[hx(r[1])]), dict(op="add_input", values=[hx(1)]),
        dict(op="select_zero", x=0, select=1), dict(op="constrain_to_constant", a=2, constant=hx(0))])
    # ---- conditionally_select_one (:145-166 selector 0 -> 1 ; :171-177 selector 1 -> value), PI = -expected
    progs["kat_select_one_sel0"] = dict(satisfied=True, program=[
        dict(op="add_input", values=[hx(r[2])]), dict(op="add_input", values=[hx(0)]),
        dict(op="select_one", y=0, select=1), dict(op="constrain_to_constant", a=2, constant=hx(0), pi=hx(-1))])
    progs["kat_select_one_sel1"] = dict(satisfied=True, program=[
        dict(op="add_input", values=[hx(r[3])]), dict(op="add_input", values=[hx(1)]),
        dict(op="select_one", y=0, select=1), dict(op="constrain_to_constant", a=2, constant=hx(0), pi=hx(-r[3]))])
    # ---- is_non_zero (:199 zero -> Err ; :205-224 mismatch -> unsatisfied ; :229-235 equal non-zero -> ok)
    progs["kat_is_non_zero_zero_errs"] = dict(satisfied=True, error="NonExistingInverse", program=[
        dict(op="add_input", values=[hx(0)]), dict(op="is_non_zero", var=0, assigned=[hx(0)])])
    progs["kat_is_non_zero_mismatch"] = dict(satisfied=False, program=[
        dict(op="add_input", values=[hx(r[4])]), dict(op="is_non_zero", var=0, assigned=[hx(r[5])])])
    progs["kat_is_non_zero_ok"] = dict(satisfied=True, program=[
        dict(op="add_input", values=[hx(r[6])]), dict(op="is_non_zero", var=0, assigned=[hx(r[6])])])
    return progs


def batch_programs():
    """Multi-instance programs: what the batched engine is compared with (row-level, restatement-pinned)."""
    progs = {}
    w = synth_wide(1, 64)
    u64s = [x & (2 ** 64 - 1) for x in synth_wide(2, 64)]
    # C1/C2 shape: range_check, [0, 2^64), k = 65; even -> uniform u64 (in range), odd -> uniform Fr
    wit = [u64s[i] if i % 2 == 0 else w[i] for i in range(9)] + [0, 1, 2 ** 63, 2 ** 64 - 1, 2 ** 64, Q - 1]
    progs["batch_range_check_k65"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in wit]), dict(op="range_check", min=hx(0), max=hx(2 ** 64), witness=0)])
    # per-instance bounds of one bit width: max-1 in [2^63, 2^64), min < max
    mx = [((x & (2 ** 63 - 1)) | 2 ** 63) + 1 for x in u64s[16:22]]
    mn = [u64s[22 + i] % mx[i] for i in range(6)]
    wt = [mn[0], mx[1] - 1, mx[2], (mn[3] - 1) % Q, (mn[4] + mx[4]) // 2, w[20]]
    progs["batch_range_check_k65_per_instance_bounds"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in wt]),
        dict(op="range_check", min=[hx(x) for x in mn], max=[hx(x) for x in mx], witness=0)])
    # C3 shape: max_bound with 252-bit bounds -> k = 253
    mx3 = [((x % 2 ** 251) | 2 ** 251) + 1 for x in w[24:29]]
    wt3 = [w[30] % mx3[0], mx3[1] - 1, mx3[2], w[31], 0]
    progs["batch_max_bound_k253"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in wt3]), dict(op="max_bound", max=[hx(x) for x in mx3], witness=0)])
    # small k, uniform bound, claims attached
    progs["batch_max_bound_k8_claims"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in (0, 99, 100, 101, 255, Q - 100)]),
        dict(op="max_bound", max=hx(101), witness=0),
        dict(op="constrain_to_constant", a=1, constant=[hx(x) for x in (1, 1, 1, 0, 0, 0)])])
    # C4 shape: is_non_zero + maybe_equal
    a = w[32:40]; b = list(a[:4]) + w[40:44]
    progs["batch_is_non_zero_maybe_equal"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in a]), dict(op="add_input", values=[hx(x) for x in b]),
        dict(op="maybe_equal", a=0, b=1),
        dict(op="is_non_zero", var=0, assigned=[hx(x) for x in a[:6]] + [hx(b[6]), hx(a[7])])])
    # is_non_zero with a zero in the middle: the loop stops there (scalar.rs:79), 1 var + 1 row already appended
    vals = [w[44], w[45], 0, w[46]]
    progs["batch_is_non_zero_error_midway"] = dict(error="NonExistingInverse", program=[
        dict(op="add_input", values=[hx(x) for x in vals]), dict(op="is_non_zero", var=0, assigned=[hx(x) for x in vals])])
    # C5 shape: mixed circuit, later calls consume earlier calls' variables
    xs = [u64s[40], w[48], u64s[41], w[49]]
    progs["batch_mixed_circuit"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in xs]),                         # 0: x
        dict(op="range_check", min=hx(0), max=hx(2 ** 64), witness=0),             # 1: in-range bit
        dict(op="select_zero", x=0, select=1),                                     # 2: x if in range else 0
        dict(op="select_one", y=0, select=1),                                      # 3: x if in range else 1
        dict(op="max_bound", max=hx(2 ** 200), witness=3),                         # 4
        dict(op="maybe_equal", a=2, b=0),                                          # 5
        dict(op="is_non_zero", var=3, assigned=[hx(x if x < 2 ** 64 else 1) for x in xs]),   # 6
        dict(op="constrain_to_constant", a=5, constant=[hx(1), hx(0), hx(1), hx(0)]),        # 7
        dict(op="constrain_to_constant", a=2, constant=hx(0), pi=[hx(-xs[0]), hx(0), hx(-xs[2]), hx(0)])])  # 8
    # dusk-plonk's native range gate (the path range.rs:9-12 recommends for power-of-two bounds): satisfied iff
    # the witness fits num_bits.  Widths cover every padding case (num_bits mod 8 = 0, 2, 4, 6) and the 256-bit maximum.
    progs["kat_range_gate_8bits_ok"] = dict(satisfied=True, program=[
        dict(op="add_input", values=[hx(200)]), dict(op="range_gate", witness=0, num_bits=8)])
    progs["kat_range_gate_8bits_too_large"] = dict(satisfied=False, program=[
        dict(op="add_input", values=[hx(256)]), dict(op="range_gate", witness=0, num_bits=8)])
    for nb in (2, 4, 6, 10, 64, 254, 256):
        top = min(2 ** nb, Q)
        vals = [0, 1, top - 1, top % Q, u64s[50] % top, w[50] % top, w[51], Q - 1]
        progs[f"batch_range_gate_{nb}bits"] = dict(program=[
            dict(op="add_input", values=[hx(x) for x in vals]), dict(op="range_gate", witness=0, num_bits=nb)])
    # range gates inside a mixed circuit: on an input column and on a gadget's output column, followed by further calls
    ys = [u64s[52] % 2 ** 32, w[52], 5, 2 ** 32]
    progs["batch_range_gate_mixed"] = dict(program=[
        dict(op="add_input", values=[hx(x) for x in ys]),                          # 0
        dict(op="range_gate", witness=0, num_bits=32),                             # 1 (no column)
        dict(op="max_bound", max=hx(2 ** 32), witness=0),                          # 2: fits-32-bits flag
        dict(op="select_zero", x=0, select=2),                                     # 3: x or 0
        dict(op="range_gate", witness=3, num_bits=34),                             # 4: always satisfied
        dict(op="range_gate", witness=2, num_bits=2),                              # 5: a bit is a quad
        dict(op="maybe_equal", a=3, b=0)])                                         # 6
    # scalar_decomposition_test of the reference (range.rs:205-233): -100 in 8 bits -> is_eq = 0
    progs["kat_decomposition_minus100_8bits"] = dict(satisfied=True, program=[
        dict(op="add_input", values=[hx(-100)]), dict(op="max_bound", max=hx(2 ** 7), witness=0)])
    return progs


def main():
    out = {}
    allp = {**kat_programs(), **batch_programs()}
    for name, spec in allp.items():
        snap = run_pymodel(spec["program"])
        exp = dict(n_rows=snap.n_rows, n_vars=snap.n_vars, unsat=snap.unsat, digest=snap.digest(),
                   results={str(k): [hx(snap.variables[v]) for v in vs] for k, vs in snap.columns.items()},
                   error=list(snap.error) if snap.error else None)
        if "satisfied" in spec:
            assert (len(snap.unsat) == 0) == spec["satisfied"], (name, snap.unsat)
        if spec.get("error"):
            assert snap.error and snap.error[1] == spec["error"], name
        else:
            assert snap.error is None, name
        out[name] = dict(program=spec["program"], expected=exp,
                         **({"satisfied": spec["satisfied"]} if "satisfied" in spec else {}))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "programs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {len(out)} programs to {path}")
    dom = domain_vectors(out)
    path = os.path.join(os.path.dirname(path), "domain.json")
    with open(path, "w") as f:
        json.dump(dom, f, indent=1, sort_keys=True)
    print(f"wrote {len(dom['fft'])} transforms and {len(dom['wire_polynomials'])} wire-polynomial sets to {path}")
    g1 = g1_vectors(out)
    path = os.path.join(os.path.dirname(path), "g1.json")
    with open(path, "w") as f:
        json.dump(g1, f, indent=1, sort_keys=True)
    print(f"wrote {len(g1['multiples'])} generator multiples, {len(g1['msm'])} sums and {len(g1['wire_commitments'])} commitment sets to {path}")


def _pt(pt):
    return None if pt is None else [hex(pt[0]), hex(pt[1])]


def g1_vectors(programs):
    """G1 / commitment fixtures from the affine big-int model of oracle/pymodel.py (second half): multiples of the
    generator, SRS powers, multi-scalar sums (with a zero scalar, a one, q-1, a repeated point and the point at infinity), and
    the four wire-polynomial commitments of two small golden programs against powers_of_g = beta^i * G."""
    from oracle import pymodel as pm
    G = pm.G1_GENERATOR
    ks = [1, 2, 3, 5, Q - 1, Q - 2] + synth_wide(97, 4)
    out = dict(multiples={hx(k): _pt(pm.g1_mul(k, G)) for k in ks}, msm={}, wire_commitments={})
    beta = synth_wide(98, 1)[0]
    out["srs"] = dict(beta=hx(beta), powers=[_pt(p) for p in pm.srs_powers(beta, 8)])
    for n in (1, 8, 40):
        pts = pm.srs_powers(synth_wide(99, 1)[0], n)
        sc = synth_wide(100 + n, n)
        sc[0] = 1
        if n >= 8:
            sc[2] = 0; sc[3] = Q - 1; pts[5] = pts[4]; pts[6] = None
        out["msm"][str(n)] = dict(points=[_pt(p) for p in pts], scalars=[hx(v) for v in sc], sum=_pt(pm.g1_msm(sc, pts)))
    for name in ("batch_is_non_zero_maybe_equal", "kat_range_check_0_ok"):
        c = run_pymodel_composer(programs[name]["program"])
        cols = pm.wire_polynomials(c)
        srs = pm.srs_powers(beta, len(cols[0]))
        out["wire_commitments"][name] = dict(log_n=pm.domain_log_size(c.n), commitments=[_pt(pm.g1_msm(col, srs)) for col in cols])
    return out


def coeff_digest(cols) -> str:
    import hashlib
    h = hashlib.sha256()
    for col in cols:
        for v in col:
            h.update(int(v).to_bytes(32, "little"))
    return h.hexdigest()


def domain_vectors(programs, max_domain: int = 512, max_programs: int = 12):
    """Evaluation-domain fixtures from the defining DFT sums of oracle/pymodel.py: subgroup generators,
    fft / ifft of seeded vectors, and the wire polynomials (ifft of the zero-padded wire columns) of small golden programs."""
    from oracle import pymodel as pm
    dom = dict(group_gen={str(k): hx(pm.group_gen(k)) for k in (0, 1, 2, 5, 16, 31, 32)}, fft={}, wire_polynomials={})
    for log_n in (0, 1, 2, 3, 6):
        vals = synth_wide(90 + log_n, 1 << log_n)
        dom["fft"][str(log_n)] = dict(input=[hx(v) for v in vals], fft=[hx(v) for v in pm.dft(vals)],
                                      ifft=[hx(v) for v in pm.dft(vals, inverse=True)])
    for name in sorted(programs):
        if len(dom["wire_polynomials"]) >= max_programs:
            break
        spec = programs[name]
        if spec["expected"]["error"] or (1 << pm.domain_log_size(spec["expected"]["n_rows"])) > max_domain:
            continue
        if name.startswith("kat_") and not name.endswith("_0_ok"):
            continue                                            # one KAT per gadget is enough here
        c = run_pymodel_composer(spec["program"])
        cols = pm.wire_polynomials(c)
        dom["wire_polynomials"][name] = dict(log_n=pm.domain_log_size(c.n), digest=coeff_digest(cols),
                                             head=[[hx(v) for v in col[:3]] for col in cols])
    return dom


def run_pymodel_composer(program):
    """The pymodel composer itself after `program` (run_pymodel returns a snapshot only)."""
    from tests import programs as P
    from oracle import pymodel as pm
    holder = {}
    orig = pm.StandardComposer

    class Capturing(orig):
        def __init__(self):
            super().__init__()
            holder["c"] = self
    pm.StandardComposer = Capturing
    try:
        P.run_pymodel(program)
    finally:
        pm.StandardComposer = orig
    return holder["c"]


if __name__ == "__main__":
    main()
