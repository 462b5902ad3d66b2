"""Gadget "programs" shared by the parity tests and the golden-vector generator.

A program is a JSON-able list of batched gadget calls over one fresh composer.  Each call is defined to be equal to
the sequential reference loop ``for i in 0..n { gadget(composer, .., operand_i) }`` (reference file:line in each
runner).  Operands that are variables are *columns*: the n variables produced by an earlier call, referenced by the
index of that call in the program.  Scalars are canonical integers written as hex strings.

Three runners execute the same program:
  * run_pymodel  -- oracle/pymodel.py   (big-int model; generates tests/golden)
  * run_oracle   -- oracle/*.c          (Montgomery limbs, reference cost structure)
  * run_engine   -- plonk_gadgets_b200  (CUDA, through the C ABI; in tests/engine_runner.py)
and return a `Snapshot` of the full composer state in the reference's own representation.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import numpy as np

Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SEL_NAMES = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_arith", "q_range", "q_logic",
             "q_fixed_group_add", "q_variable_group_add")


def hx(v: int) -> str:
    return hex(v % Q)


def unhx(s) -> int:
    return int(s, 16) if isinstance(s, str) else int(s)


@dataclass
class Snapshot:
    """Composer state after a program, canonical integers everywhere."""
    n_rows: int
    n_vars: int
    variables: list            # n_vars canonical ints
    wires: np.ndarray          # (4, n_rows) uint64
    selectors: list            # 11 lists of n_rows canonical ints
    dense_pi: list             # n_rows canonical ints
    unsat: list                # unsatisfied row indices (verdict: satisfied iff empty)
    columns: dict = field(default_factory=dict)   # op index -> list of Variable ids returned by that call
    error: tuple | None = None  # (op index, "NonExistingInverse", completed instances)

    def digest(self) -> str:
        h = hashlib.sha256()
        h.update(self.n_vars.to_bytes(8, "little"))
        h.update(self.n_rows.to_bytes(8, "little"))
        for v in self.variables:
            h.update(int(v).to_bytes(32, "little"))
        for i in range(self.n_rows):
            for w in range(4):
                h.update(int(self.wires[w, i]).to_bytes(8, "little"))
            for k in range(len(SEL_NAMES)):
                h.update(int(self.selectors[k][i]).to_bytes(32, "little"))
            h.update(int(self.dense_pi[i]).to_bytes(32, "little"))
        return h.hexdigest()

    def results(self, op_index: int) -> list:
        """Canonical values of the variables a call returned."""
        return [self.variables[v] for v in self.columns[op_index]]


def _vals(op, key):
    v = op[key]
    if isinstance(v, (list, tuple)):
        return [unhx(x) for x in v]
    return [unhx(v)]


# ------------------------------------------------------------------------------------------------ pymodel
def run_pymodel(program) -> Snapshot:
    from oracle import pymodel as pm
    c = pm.StandardComposer()
    cols, error = {}, None
    for idx, op in enumerate(program):
        kind = op["op"]
        if kind == "add_input":                       # allocated_scalar.rs:27-30 / composer.add_input
            cols[idx] = [c.add_input(v) for v in _vals(op, "values")]
        elif kind == "range_check":                   # range.rs:27-43
            wit = cols[op["witness"]]; mn = _vals(op, "min"); mx = _vals(op, "max")
            cols[idx] = [pm.range_check(c, mn[i % len(mn)], mx[i % len(mx)], pm.AllocatedScalar(w, c.variables[w]))
                         for i, w in enumerate(wit)]
        elif kind == "max_bound":                     # range.rs:82-113
            wit = cols[op["witness"]]; mx = _vals(op, "max")
            cols[idx] = [pm.max_bound(c, mx[i % len(mx)], pm.AllocatedScalar(w, c.variables[w]))[0] for i, w in enumerate(wit)]
        elif kind == "maybe_equal":                   # scalar.rs:105-140
            cols[idx] = [pm.maybe_equal(c, pm.AllocatedScalar(a, c.variables[a]), pm.AllocatedScalar(b, c.variables[b]))
                         for a, b in zip(cols[op["a"]], cols[op["b"]])]
        elif kind == "is_non_zero":                   # scalar.rs:63-97, loop with `?`
            vs = cols[op["var"]]; asg = _vals(op, "assigned")
            for i, v in enumerate(vs):
                try:
                    pm.is_non_zero(c, v, asg[i])
                except pm.NonExistingInverse:
                    error = (idx, "NonExistingInverse", i)
                    break
            if error:
                break
        elif kind == "select_zero":                   # scalar.rs:21-27
            cols[idx] = [pm.conditionally_select_zero(c, x, s) for x, s in zip(cols[op["x"]], cols[op["select"]])]
        elif kind == "select_one":                    # scalar.rs:36-59
            cols[idx] = [pm.conditionally_select_one(c, y, s) for y, s in zip(cols[op["y"]], cols[op["select"]])]
        elif kind == "range_gate":                    # composer.range_gate [dusk-plonk], recommended at range.rs:9-12
            for w in cols[op["witness"]]:
                c.range_gate(w, int(op["num_bits"]))
        elif kind == "constrain_to_constant":         # tests/range_gadgets_tests.rs:26, tests/scalar_gadgets_tests.rs:135
            a = cols[op["a"]]; k = _vals(op, "constant"); pi = _vals(op, "pi") if op.get("pi") is not None else None
            for i, v in enumerate(a):
                c.constrain_to_constant(v, k[i % len(k)], None if pi is None else pi[i % len(pi)])
        else:
            raise ValueError(kind)
    wires = np.array([c.w_l, c.w_r, c.w_o, c.w_4], dtype=np.uint64).reshape(4, c.n)
    return Snapshot(c.n, len(c.variables), list(c.variables), wires, [list(c.sel[k]) for k in SEL_NAMES],
                    c.construct_dense_pi_vec(), c.unsatisfied_rows(), cols, error)


# ------------------------------------------------------------------------------------------------ C oracle
def snapshot_of_oracle(c, cols=None, error=None) -> Snapshot:
    from oracle import binding as ob
    n = c.n
    sel = c.selectors()
    bad, _first = c.check()
    snap = Snapshot(n, c.n_vars, ob.to_ints(c.variables()), c.wires(), [ob.to_ints(sel[k]) for k in range(len(SEL_NAMES))],
                    ob.to_ints(c.dense_pi()), [], cols or {}, error)
    # recompute the list of unsatisfied rows from the dump (big-int), and cross-check the oracle's own count
    snap.unsat = unsat_rows(snap)
    assert len(snap.unsat) == bad, (len(snap.unsat), bad)
    return snap


def run_oracle(program, return_composer: bool = False):
    from oracle import binding as ob
    c = ob.Composer()
    cols, error = {}, None

    def bounds(op, key, n):
        v = _vals(op, key)
        assert len(v) in (1, n)
        return ob.from_ints(v)

    for idx, op in enumerate(program):
        kind = op["op"]
        if kind == "add_input":
            cols[idx] = c.add_input_batch(ob.from_ints(_vals(op, "values")))
        elif kind == "range_check":
            wit = cols[op["witness"]]
            cols[idx] = c.range_check_batch(bounds(op, "min", len(wit)), bounds(op, "max", len(wit)), wit)
        elif kind == "max_bound":
            wit = cols[op["witness"]]
            cols[idx], _k = c.max_bound_batch(bounds(op, "max", len(wit)), wit)
        elif kind == "maybe_equal":
            cols[idx] = c.maybe_equal_batch(cols[op["a"]], cols[op["b"]])
        elif kind == "is_non_zero":
            e, done = c.is_non_zero_batch(cols[op["var"]], ob.from_ints(_vals(op, "assigned")))
            if e:
                assert e == ob.ERR_NON_EXISTING_INVERSE
                error = (idx, "NonExistingInverse", done)
                break
        elif kind == "select_zero":
            cols[idx] = c.select_zero_batch(cols[op["x"]], cols[op["select"]])
        elif kind == "select_one":
            cols[idx] = c.select_one_batch(cols[op["y"]], cols[op["select"]])
        elif kind == "range_gate":
            c.range_gate_batch(cols[op["witness"]], int(op["num_bits"]))
        elif kind == "constrain_to_constant":
            a = cols[op["a"]]
            pi = bounds(op, "pi", len(a)) if op.get("pi") is not None else None
            c.constrain_to_constant_batch(a, bounds(op, "constant", len(a)), pi)
        else:
            raise ValueError(kind)
    snap = snapshot_of_oracle(c, {k: [int(x) for x in v] for k, v in cols.items()}, error)
    return (snap, c) if return_composer else snap


def unsat_rows(s: Snapshot) -> list:
    """Gate equation q_arith*(q_m*a*b + q_l*a + q_r*b + q_o*c + q_4*d + PI + q_c) + q_range*(delta(c-4d) + delta(b-4c) +
    delta(a-4b) + delta(d_next-4a)), delta(f) = f(f-1)(f-2)(f-3), evaluated with Python ints."""
    sel = {k: s.selectors[i] for i, k in enumerate(SEL_NAMES)}
    out = []

    def delta(f):
        return f * (f - 1) * (f - 2) * (f - 3)
    for i in range(s.n_rows):
        a, b, c, d = (s.variables[int(s.wires[w, i])] for w in range(4))
        g = sel["q_arith"][i] * (sel["q_m"][i] * a * b + sel["q_l"][i] * a + sel["q_r"][i] * b + sel["q_o"][i] * c
                                 + sel["q_4"][i] * d + s.dense_pi[i] + sel["q_c"][i])
        if sel["q_range"][i]:
            d_next = s.variables[int(s.wires[3, (i + 1) % s.n_rows])]
            g += sel["q_range"][i] * (delta(c - 4 * d) + delta(b - 4 * c) + delta(a - 4 * b) + delta(d_next - 4 * a))
        g %= Q
        if g:
            out.append(i)
    return out


# ------------------------------------------------------------------------------------------------ synthetic scalars
def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (counter-based: output j depends only on seed + j)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


SEED = 0x706C6F6E6B5F6732


def synth_wide(stream: int, n: int) -> list:
    """n uniform Fr values: 512-bit SplitMix64 draws reduced mod q (canonical ints)."""
    words = splitmix64(SEED ^ (stream * 0xD1342543DE82EF95 & (2 ** 64 - 1)), 8 * n).reshape(n, 8)
    out = []
    for row in words:
        v = 0
        for j in range(8):
            v |= int(row[j]) << (64 * j)
        out.append(v % Q)
    return out


def expected_sigma(oracle_composer) -> np.ndarray:
    """Copy-constraint cycles from the oracle's perm.variable_map: (4, n_rows) uint64, sigma[w, r] = successor (row*4 + wire)
    of wire position (r, w) in the cycle of its Variable (positions in insertion order, last wraps to first)."""
    n = oracle_composer.n
    # every position starts as its own successor (compute_sigma_permutations initialises sigma with the identity); range_gate
    # leaves three wire positions of its last gate out of the map
    sigma = (np.arange(n, dtype=np.uint64)[None, :] * np.uint64(4) + np.arange(4, dtype=np.uint64)[:, None]).astype(np.uint64)
    seen = 0
    for v in range(oracle_composer.n_vars):
        uses = oracle_composer.perm_of(v)
        for k, (r, w) in enumerate(uses):
            nr, nw = uses[(k + 1) % len(uses)]
            sigma[w, r] = nr * 4 + nw
        seen += len(uses)
    assert seen <= 4 * n and (4 * n - seen) % 3 == 0
    return sigma
