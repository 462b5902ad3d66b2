"""The C oracle against (i) the reference's own known-answer tests (verdict level) and (ii) the golden vectors
generated from the independent big-int model (row level: every variable, wire index, selector and public input)."""
import pytest

from tests.programs import run_oracle, run_pymodel, synth_wide, hx, Q


def test_golden_file_is_current(golden):
    """tests/golden/programs.json is what `python -m oracle.gen_golden` produces from oracle/pymodel.py."""
    for name, spec in golden.items():
        snap = run_pymodel(spec["program"])
        assert snap.digest() == spec["expected"]["digest"], name


def test_oracle_matches_golden(golden):
    assert len(golden) >= 40
    for name, spec in golden.items():
        exp = spec["expected"]
        snap = run_oracle(spec["program"])
        assert (snap.n_rows, snap.n_vars) == (exp["n_rows"], exp["n_vars"]), name
        assert snap.unsat == exp["unsat"], name
        assert snap.digest() == exp["digest"], name
        assert (list(snap.error) if snap.error else None) == exp["error"], name
        for k, vals in exp["results"].items():
            assert [hx(v) for v in snap.results(int(k))] == vals, (name, k)


def test_reference_kats_verdicts(golden):
    """Reference KATs: the claimed outcome is satisfiable, the negated claim is not
    (/root/reference/tests/range_gadgets_tests.rs:57-78,:120-169; tests/scalar_gadgets_tests.rs:36-235)."""
    n = 0
    for name, spec in golden.items():
        if not name.startswith("kat_"):
            continue
        snap = run_oracle(spec["program"])
        assert (len(snap.unsat) == 0) == spec["satisfied"], name
        n += 1
    assert n >= 36


def test_row_and_variable_counts():
    """Closed forms of SURVEY.md section 3: range_check 4k+11 rows / 2k+523 vars; max_bound 2k+5 / k+261."""
    for bound_bits in (1, 7, 18, 64, 128, 200, 252, 253):
        k = bound_bits + 1
        p = [dict(op="add_input", values=[hx(5)]), dict(op="range_check", min=hx(0), max=hx(2 ** bound_bits), witness=0)]
        s = run_oracle(p)
        assert s.n_rows == 3 + 4 * k + 11 and s.n_vars == 5 + 1 + 2 * k + 523
        p = [dict(op="add_input", values=[hx(5)]), dict(op="max_bound", max=hx(2 ** bound_bits), witness=0)]
        s = run_oracle(p)
        assert s.n_rows == 3 + 2 * k + 5 and s.n_vars == 5 + 1 + k + 261


@pytest.mark.parametrize("bits", [2, 9, 33, 64, 65, 127, 252])
def test_boundary_witnesses(bits):
    """min, max-1 in range; max, min-1 out of range (when the difference does not wrap into k bits)."""
    r = synth_wide(7, 2)
    mx = (r[0] % 2 ** (bits - 1)) | 2 ** (bits - 1)
    mx += 1                                      # max-1 has exactly `bits` bits
    mn = r[1] % mx
    wit = [mn, mx - 1, mx, (mn - 1) % Q, 0, Q - 1]
    p = [dict(op="add_input", values=[hx(x) for x in wit]), dict(op="range_check", min=hx(mn), max=hx(mx), witness=0)]
    so, sp = run_oracle(p), run_pymodel(p)
    assert so.digest() == sp.digest() and so.unsat == []
    res = so.results(1)
    assert res[0] == 1 and res[1] == 1 and res[2] == 0
    if mn > 0:
        assert res[3] == 0
    assert res[4] == (1 if mn == 0 else 0)
