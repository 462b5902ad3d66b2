"""Fault injection into the range gadgets' own segments, shared by the CPU (tests/emu) and GPU parity tests.

The gadgets only ever produce consistent witnesses, so a check kernel that returned "0 unsatisfied rows" unconditionally would
pass every positive test.  Here stored variables of a *large* range_check / max_bound segment are overwritten
(pg_poke_variable: V, accumulators A_j, U, Zv, Y of both decompositions, and the final product O -- the Variables of
/root/reference/src/range.rs:93-102,:138-155,:42 and /root/reference/src/scalar.rs:111-126) in instances at both ends of the
batch, at block/warp boundaries and in the middle, and pg_check must report exactly the rows a big-int evaluation of the
engine's own row dump reports.  With n >= 320 * SMs instances the verdict comes from the one-thread-per-instance kernels
(`k_check<GENERIC>` / the compiled row program `k_check_prog`), which the test asserts through pg_get_check_stats.
"""
from __future__ import annotations

import random

import numpy as np

import plonk_gadgets_b200 as pg
from tests.programs import Q, synth_wide

SEL = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c")


def eval_rows(ob, rows) -> list:
    """Unsatisfied local row indices of a materialised row range (arithmetic widget), evaluated with Python ints."""
    w = [ob.to_ints(rows["w_val"][k]) for k in range(4)]
    s = [ob.to_ints(rows["sel"][k]) for k in range(6)]
    pi = ob.to_ints(rows["pi"])
    out = []
    for r in range(len(pi)):
        a, b, c, d = (w[k][r] for k in range(4))
        g = s[0][r] * a * b + s[1][r] * a + s[2][r] * b + s[3][r] * c + s[4][r] * d + s[5][r] + pi[r]
        if g % Q:
            out.append(r)
    return out


def decomposition_targets(k: int, base_row: int) -> dict:
    """(local row, wire) whose Variable is the named one, for a decomposition D(v, k) preceded by its V row at `base_row`
    (row templates of SURVEY.md 8a: V row, A_0 row, k x (boolean, accumulate), u row, y row, y*u row)."""
    t = {"V": (base_row, 2), "A0": (base_row + 1, 0)}
    for j in (1, 2, k // 2, k - 1, k):
        if 1 <= j <= k:
            t[f"A{j}"] = (base_row + 2 * j + 1, 2)       # output wire of the j-th accumulate row
    t["U"] = (base_row + 2 * k + 2, 2)
    t["Zv"] = (base_row + 2 * k + 3, 0)
    t["Y"] = (base_row + 2 * k + 3, 2)
    return t


def poked_range_segment(make_composer, ob, n: int, gadget: str = "range_check", bits: int = 64, per_instance_bounds: bool = False,
                        modes=(pg.CHECK_GENERIC, pg.CHECK_SPARSE), expect_kind: dict | None = None, seed: int = 5, wit_dev=None, fused: bool = False):
    """`expect_kind`: {mode: kind name of pg_get_check_stats} that must have evaluated the segment's rows (None: not asserted).
    `fused` (PG_F_FUSED_CHECK, structure-aware mode): the rows are first evaluated inside witness generation (kind "fused", no check
    launch for the segment); a poke sends the segment back to the check kernel, which must then find the fault."""
    rng = random.Random(seed)
    mx_i = 2 ** bits
    k = bits + 1
    rows_per = 4 * k + 11 if gadget == "range_check" else 2 * k + 5
    vars_per = 2 * k + 523 if gadget == "range_check" else k + 261
    vals = synth_wide(900 + bits, min(n, 4096))
    wit = [(vals[i % len(vals)] % mx_i) if i % 2 == 0 else vals[i % len(vals)] for i in range(n)]
    targets = decomposition_targets(k, 0)
    if gadget == "range_check":
        targets.update({name + "'": pos for name, pos in decomposition_targets(k, 2 * k + 5).items()})
        targets["O"] = (4 * k + 10, 2)
    instances = sorted({0, 1, 31, 32, 127, 128, n // 2 + 5, n - 129, n - 2, n - 1} & set(range(n)))
    for mode in modes:
        c = make_composer(check_mode=mode, fused_check=True) if fused else make_composer(check_mode=mode)
        c.check_stats(reset=True)
        w = c.add_input(ob.from_ints(wit))
        if per_instance_bounds:
            mxs = ob.from_ints([mx_i - (i % 7) for i in range(n)])       # bitlen(max - 1) stays `bits`
            mns = ob.from_ints([i % 5 for i in range(n)])
        else:
            mxs, mns = ob.from_ints([mx_i]), ob.from_ints([0])
        y = pg.range_check(c, mns, mxs, w) if gadget == "range_check" else pg.max_bound(c, mxs, w)[0]
        base_row, base_var = 3, 5 + n
        assert c.circuit_size() == base_row + rows_per * n and c.num_variables() == base_var + vars_per * n
        if fused:                                           # evaluated while generated; pg_check launches for the 3 fresh rows only
            assert mode == pg.CHECK_SPARSE
            assert c.check_circuit_satisfied() == (0, None)
            stats = c.check_stats(reset=True)
            assert stats["fused"][1] == rows_per * n and stats["program"][1] == 0 and stats["instance_generic"][1] == 0, stats
            assert c.check_sharded(None, 0) == (0, None, 0)  # the sharded verdict takes the recorded result too (no launch for the segment)
            assert c.check_stats(reset=True)["program"][1] == 0
        else:
            c.check_stats(reset=True)
            assert c.check_circuit_satisfied() == (0, None)
            stats = c.check_stats(reset=True)
            if expect_kind and expect_kind.get(mode):
                assert stats[expect_kind[mode]][1] >= rows_per * n, (mode, stats)
        # single pokes: every named Variable, in rotating instances
        names = sorted(targets)
        for t_i, name in enumerate(names):
            i = instances[t_i % len(instances)]
            if name.startswith("Zv") and i % 2 == 0:       # in-range instances have u = 0: y = 1 - z*u holds for any z
                i = i + 1 if i + 1 < n else i - 1
            r_loc, wire = targets[name]
            w_idx = c.rows(base_row + rows_per * i, rows_per, want=("w_idx",))["w_idx"]
            var = int(w_idx[wire, r_loc])
            assert base_var + vars_per * i <= var < base_var + vars_per * (i + 1), (name, var)
            old = c.variables(var, 1)[0].copy()
            new = (ob.to_ints(old[None])[0] + rng.choice([1, 2, Q - 1, rng.randrange(1, Q)])) % Q
            c.poke_variable(var, ob.from_ints([new])[0])
            exp = [base_row + rows_per * i + r for r in eval_rows(ob, c.rows(base_row + rows_per * i, rows_per, want=("w_val", "sel", "pi")))]
            assert exp and base_row + rows_per * i + r_loc in exp, (name, i, exp)
            got = c.check_circuit_satisfied()
            assert got == (len(exp), exp[0]), (mode, name, i, got, exp)
            c.poke_variable(var, old)
            assert c.check_circuit_satisfied() == (0, None), (mode, name, i)
        # many faults at once, spread over the batch (both ends, block boundaries, the middle)
        spread = sorted({0, 1, 127, 128, n // 3, n // 2, n // 2 + 1, n - 129, n - 1} & set(range(n)))
        all_exp, undo = [], []
        for j, i in enumerate(spread):
            name = names[(3 * j + 1) % len(names)]
            if name.startswith("Zv"):
                name = "U"
            r_loc, wire = targets[name]
            var = int(c.rows(base_row + rows_per * i + r_loc, 1, want=("w_idx",))["w_idx"][wire, 0])
            old = c.variables(var, 1)[0].copy()
            c.poke_variable(var, ob.from_ints([(ob.to_ints(old[None])[0] + 1 + j) % Q])[0])
            undo.append((var, old))
        for i in spread:
            all_exp += [base_row + rows_per * i + r for r in eval_rows(ob, c.rows(base_row + rows_per * i, rows_per, want=("w_val", "sel", "pi")))]
        all_exp = sorted(set(all_exp))
        assert len(all_exp) >= len(spread)
        c.check_stats(reset=True)
        assert c.check_circuit_satisfied() == (len(all_exp), all_exp[0]), (mode, all_exp[:6])
        stats = c.check_stats(reset=True)
        if expect_kind and expect_kind.get(mode):
            assert stats[expect_kind[mode]][1] >= rows_per * n, (mode, stats)
        for var, old in undo:
            c.poke_variable(var, old)
        assert c.check_circuit_satisfied() == (0, None)
        # the result column still reads what the gadget produced
        res = ob.to_ints(y.values(0, min(n, 64)))
        assert all(v in (0, 1) for v in res)
        c.close()


# ---------------------------------------------------------------------------------------------------- is_non_zero, every Result kept
def non_zero_flags_vs_oracle(make_composer, ob, n: int = 300, zero_every: int = 37, mismatch_every: int = 53, modes=(pg.CHECK_GENERIC, pg.CHECK_SPARSE)):
    """pg_is_non_zero_batch_flags against the oracle run as the reference loop WITHOUT `?`:
    ``for i: results[i] = is_non_zero(composer, var_i, assigned_i)`` (/root/reference/src/scalar.rs:63-97; an Err leaves the
    1 variable + 1 row of :69-71 behind and the loop goes on).  PG_NZ_REFERENCE must give that composer bit for bit; PG_NZ_UNIFORM
    gives 3 variables + 3 rows everywhere with inv = 0 and an unsatisfied last row for the errored instances."""
    from tests.programs import snapshot_of_oracle
    from tests.engine_runner import snapshot_of_engine
    vals = [v or 1 for v in synth_wide(700, n)]
    assigned = list(vals)
    zeros = [i for i in range(n) if i % zero_every == zero_every - 1] + ([0, 1] if n > 4 else [])      # adjacent errors, error at instance 0
    for i in zeros:
        assigned[i] = 0
    mism = [i for i in range(n) if i % mismatch_every == 3 and i not in zeros]
    for i in mism:
        assigned[i] = (assigned[i] + 7) % Q or 1
    zeros = sorted(set(zeros))
    # oracle: one call per instance, errors ignored
    oc = ob.Composer()
    ov = oc.add_input_batch(ob.from_ints(vals))
    o_flags = []
    for i in range(n):
        e, _done = oc.is_non_zero_batch(ov[i:i + 1], ob.from_ints([assigned[i]]))
        o_flags.append(1 if e else 0)
    assert [i for i, f in enumerate(o_flags) if f] == zeros
    so = snapshot_of_oracle(oc)
    for mode in modes:
        c = make_composer(check_mode=mode)
        v = c.add_input(ob.from_ints(vals))
        flags = pg.is_non_zero_flags(c, v, ob.from_ints(assigned), layout=pg.NZ_REFERENCE)
        assert flags.tolist() == o_flags and c.last_n_err == len(zeros)
        se = snapshot_of_engine(c, {}, None, ob)
        assert (se.n_rows, se.n_vars) == (so.n_rows, so.n_vars) == (3 + 3 * n - 2 * len(zeros), 5 + n + 3 * n - 2 * len(zeros))
        assert se.digest() == so.digest() and se.unsat == so.unsat
        assert len(se.unsat) == 2 * len(mism) + len(zeros)         # a wrong value_assigned breaks assert_equal and var*inv = 1; a zero one assert_equal
        from tests.programs import expected_sigma
        assert (c.permutation() == expected_sigma(oc)).all()
        c.close()
        # uniform layout
        c = make_composer(check_mode=mode)
        v = c.add_input(ob.from_ints(vals))
        flags = pg.is_non_zero_flags(c, v, ob.from_ints(assigned), layout=pg.NZ_UNIFORM)
        assert flags.tolist() == o_flags and c.last_n_err == len(zeros)
        assert c.circuit_size() == 3 + 3 * n and c.num_variables() == 5 + 4 * n
        var = ob.to_ints(c.variables(5 + n, 3 * n))
        for i in range(n):
            va, inv, one = var[3 * i: 3 * i + 3]
            assert va == assigned[i] and one == 1 and inv == (pow(va, -1, Q) if va else 0), i
        exp = sorted([3 + 3 * i + 2 for i in zeros] + [3 + 3 * i + r for i in mism for r in (0, 2)]
                     + [3 + 3 * i for i in zeros])                 # var != 0 = var_assigned: assert_equal fails as well
        assert c.check_circuit_satisfied() == (len(exp), exp[0])
        c.close()
    # no zero at all: both layouts are the plain batch
    c = make_composer()
    v = c.add_input(ob.from_ints(vals))
    assert not pg.is_non_zero_flags(c, v, ob.from_ints(vals), layout=pg.NZ_REFERENCE).any() and c.last_n_err == 0
    assert c.circuit_size() == 3 + 3 * n and c.check_circuit_satisfied() == (0, None)
    c.close()


def unreduced_inputs_rejected(make_composer, ob):
    """Scalars >= q cannot be BlsScalars: single ones are refused on the spot, batches by the next call that reads the device
    counters (PG_ERR_ARG = -2), and nothing hangs (a table value equal to q used to make the block inversion spin)."""
    import pytest
    q_limbs = np.frombuffer(Q.to_bytes(32, "little"), dtype=np.uint64).reshape(1, 4)
    big = np.full((1, 4), 2 ** 64 - 1, dtype=np.uint64)
    good = ob.from_ints([5, 6, 7, 8])
    # a batch with one unreduced value: reported by the verdict
    c = make_composer()
    bad = good.copy(); bad[2] = q_limbs[0]
    w = c.add_input(bad)
    with pytest.raises(pg.EngineError) as e:
        c.check_circuit_satisfied()
    assert e.value.code == -2 and "not fully reduced" in str(e.value) and "index 2" in str(e.value)
    c.reset()                                                         # the composer is usable again
    w = c.add_input(good)
    pg.maybe_equal(c, w, w)
    assert c.check_circuit_satisfied() == (0, None)
    # ... by pg_sync, and the inversion-based gadgets terminate on such a table
    c.reset()
    w = c.add_input(bad)
    pg.maybe_equal(c, w, c.add_input(good))
    pg.range_check(c, ob.from_ints([0]), ob.from_ints([2 ** 64]), w)
    with pytest.raises(pg.EngineError) as e:
        c.sync()
    assert e.value.code == -2
    # value_assigned, per-instance bounds, per-instance constants
    c.reset()
    w = c.add_input(good)
    with pytest.raises(pg.EngineError) as e:
        pg.is_non_zero(c, w, bad)
    assert e.value.code == -2
    c.reset(); w = c.add_input(good)
    with pytest.raises(pg.EngineError) as e:
        pg.range_check(c, ob.from_ints([0] * 4), np.concatenate([ob.from_ints([2 ** 10] * 3), big]), w)
    assert e.value.code == -2
    c.reset(); w = c.add_input(good)
    c.constrain_to_constant(w, bad)
    with pytest.raises(pg.EngineError) as e:
        c.check_circuit_satisfied()
    assert e.value.code == -2
    # single scalars: immediately
    c.reset(); w = c.add_input(good)
    for call in (lambda: pg.range_check(c, ob.from_ints([0]), q_limbs, w), lambda: pg.max_bound(c, big, w),
                 lambda: c.constrain_to_constant(w, q_limbs), lambda: c.poke_variable(5, big[0])):
        with pytest.raises(pg.EngineError) as e:
            call()
        assert e.value.code == -2
    assert c.check_circuit_satisfied()[0] == 0
    c.close()


# ---------------------------------------------------------------------------------------------------- fused check, scalar gadgets
def fused_scalar_gadgets(make_composer, ob, n: int = 90):
    """PG_F_FUSED_CHECK over maybe_equal / is_non_zero (every Result kept) / conditionally_select_*: the rows are evaluated by the
    gadgets' own kernels (kind "fused": the check launches only for the fresh rows and the claim rows), the verdict equals the one of
    an unfused structure-aware composer on the same inputs -- including the rows errored and mismatching is_non_zero instances fail --
    and overwriting a Variable afterwards sends the affected segments back to the check kernel."""
    vals = synth_wide(321, n)
    other = [v if i % 2 == 0 else (v + 1 + i) % Q for i, v in enumerate(vals)]
    assigned = list(vals)
    assigned[7] = 0; assigned[n - 2] = 0                    # Err(NonExistingInverse): rows 0 and 2 of those instances are unsatisfied
    assigned[11] = (assigned[11] + 5) % Q                   # mismatch: rows 0 and 2 as well
    sel = [v & 1 for v in synth_wide(322, n)]
    verdicts = []
    for fused in (False, True):
        c = make_composer(check_mode=pg.CHECK_SPARSE, fused_check=fused)
        c.check_stats(reset=True)
        a = c.add_input(ob.from_ints(vals)); b = c.add_input(ob.from_ints(other)); s = c.add_input(ob.from_ints(sel))
        eq = pg.maybe_equal(c, a, b)
        nz_first_var = c.num_variables()
        flags = pg.is_non_zero_flags(c, a, ob.from_ints(assigned), pg.NZ_UNIFORM)
        so = pg.conditionally_select_one(c, a, s)
        sz = pg.conditionally_select_zero(c, so, s)
        c.constrain_to_constant(eq, ob.from_ints([1 if i % 2 == 0 else 0 for i in range(n)]))        # true claims
        assert flags.sum() == 2 and flags[7] and flags[n - 2]
        got = c.check_circuit_satisfied()
        stats = c.check_stats(reset=True)
        n_rows = c.circuit_size()
        exp = eval_rows(ob, c.rows(0, n_rows, want=("w_val", "sel", "pi")))
        assert got == (len(exp), exp[0]) and len(exp) == 6, (fused, got, exp)
        if fused:
            assert stats["fused"][1] == n * (3 + 3 + 4 + 1), stats            # maybe_equal, is_non_zero, select_one, select_zero
            assert stats["program"][1] + stats["rowpar"][1] + stats["instance_terms"][1] == 3 + n, stats    # fresh rows + claim rows only
        verdicts.append(got)
        assert ob.to_ints(sz.values()) == [v * s_ * s_ % Q if s_ else 0 for v, s_ in zip(vals, sel)]
        # a poke after generation: the segment (and the later ones) go back to the kernel, which finds exactly the big-int verdict
        first_var, stride = c._col_geometry(so.col)
        var = first_var + stride * (n // 2)
        old = c.variables(var, 1)[0].copy()
        c.poke_variable(var, ob.from_ints([(ob.to_ints(old[None])[0] + 3) % Q])[0])
        exp2 = eval_rows(ob, c.rows(0, n_rows, want=("w_val", "sel", "pi")))
        assert len(exp2) > len(exp) and c.check_circuit_satisfied() == (len(exp2), exp2[0])
        c.poke_variable(var, old)
        assert c.check_circuit_satisfied() == got
        # a poke into the segment whose generation-time verdict held unsatisfied rows (the errored / mismatching is_non_zero instances)
        # while an earlier segment (maybe_equal) had been verified at generation too: those rows must not be counted twice
        var = nz_first_var + 3 * (n // 3) + 1                   # the inverse of instance n // 3 (VA, I, ONE per instance)
        old = c.variables(var, 1)[0].copy()
        c.poke_variable(var, ob.from_ints([(ob.to_ints(old[None])[0] + 9) % Q])[0])
        exp3 = eval_rows(ob, c.rows(0, n_rows, want=("w_val", "sel", "pi")))
        assert len(exp3) == len(exp) + 1 and c.check_circuit_satisfied() == (len(exp3), exp3[0])
        c.poke_variable(var, old)
        assert c.check_circuit_satisfied() == got
        # gadgets called after a poke are recorded afresh
        y2 = pg.maybe_equal(c, a, a)
        c.constrain_to_constant(y2, ob.from_ints([1] * n))
        assert c.check_circuit_satisfied() == got
        c.close()
    assert verdicts[0] == verdicts[1]
