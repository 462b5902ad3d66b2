"""Independent big-integer model of the plonk_gadgets hot path (TEST INFRASTRUCTURE ONLY).

This is the *second opinion* for the C oracle in this directory and the generator of the
golden vectors under ``tests/golden``.  It is deliberately written with plain Python ints
(canonical residues mod q, no Montgomery form, no limbs) so that it shares no arithmetic
code with ``oracle/*.c`` or with the CUDA kernels.

PARITY STATUS: the reference crate (Rust) and its dependency ``dusk-plonk 0.8`` cannot be
built in this environment and the reference holds no row-level golden vectors, so parity is
*pinned at verdict level* by the reference's own known-answer tests (replayed in
``tests/test_oracle_kats.py``) and **unpinned at row level** (wire/selector/variable dumps are
pinned only by this restatement).

What it follows (reference file:line):
  * gadgets                : /root/reference/src/range.rs:21-189, /root/reference/src/scalar.rs:21-140
  * AllocatedScalar        : /root/reference/src/allocated_scalar.rs:17-31
  * Error                  : /root/reference/src/errors.rs:13-18
  * composer / Fr semantics: third-party ``dusk-plonk = "0.8"`` / ``dusk-bls12_381`` (not vendored in
    /root/reference); restated from their published behaviour as recorded in SURVEY.md Appendix A.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R = (1 << 256) % Q          # Montgomery radix residue (BlsScalar::one() raw limbs)
R2 = (R * R) % Q
R3 = (R2 * R) % Q

SELECTORS = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_arith", "q_range", "q_logic",
             "q_fixed_group_add", "q_variable_group_add")


class NonExistingInverse(Exception):
    """errors.rs:17 -- the only error the gadgets can return (scalar.rs:79)."""


# --------------------------------------------------------------------------- Fr helpers
def fr(x: int) -> int:
    return x % Q


def inv_or_none(x: int):
    """BlsScalar::invert(): None iff x == 0, else the unique inverse."""
    x %= Q
    if x == 0:
        return None
    return pow(x, Q - 2, Q)


def to_mont_limbs(x: int):
    """Raw BlsScalar representation: 4 little-endian u64 limbs of x*2^256 mod q."""
    m = (x % Q) * R % Q
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def to_bytes(x: int) -> bytes:
    """BlsScalar::to_bytes(): canonical little-endian 32 bytes."""
    return (x % Q).to_bytes(32, "little")


def mont_bytes(x: int) -> bytes:
    """In-memory bytes of a BlsScalar ([u64;4] LE limbs of the Montgomery form)."""
    return ((x % Q) * R % Q).to_bytes(32, "little")


def from_bytes_wide(b: bytes) -> int:
    """BlsScalar::from_bytes_wide: 64 LE bytes reduced mod q (d0*R2 + d1*R3 in Montgomery terms)."""
    assert len(b) == 64
    return int.from_bytes(b, "little") % Q


def bits_count(s: int) -> int:
    """range.rs:173-181: reduce(); counter=1; while s > 1 { s >>= 1; counter += 1 }."""
    s %= Q
    counter = 1
    while s > 1:
        s >>= 1
        counter += 1
    return counter


def num_bits_closest_power_of_two(s: int) -> int:
    """range.rs:185-189: bits_count(2^bits_count(s) mod q)."""
    n = bits_count(s)
    return bits_count(pow(2, n, Q))


def scalar_to_bits(s: int):
    """range.rs:161-170: bit[8j+i] = (to_bytes()[j] >> i) & 1 -> LSB-first bits of the canonical integer."""
    by = to_bytes(s)
    out = []
    for byte in by:
        for i in range(8):
            out.append((byte >> i) & 1)
    return out


# --------------------------------------------------------------------------- composer
@dataclass(frozen=True)
class AllocatedScalar:
    """allocated_scalar.rs:17-23."""
    var: int
    scalar: int

    @staticmethod
    def allocate(composer: "StandardComposer", scalar: int) -> "AllocatedScalar":
        """allocated_scalar.rs:27-30."""
        return AllocatedScalar(composer.add_input(scalar), scalar % Q)


class StandardComposer:
    """Arithmetic-row subset of dusk-plonk 0.8's StandardComposer."""

    def __init__(self):
        self.n = 0
        self.variables = []                       # Variable(i) -> canonical value
        self.w_l, self.w_r, self.w_o, self.w_4 = [], [], [], []
        self.sel = {name: [] for name in SELECTORS}
        self.public_inputs = {}                   # sparse: row -> value
        self.perm_map = {}                        # perm.variable_map: Variable -> [(row, wire)] in insertion order
        # new(): zero_var (placeholder Variable(0), then bound by its own constant row) and the two dummy constraints
        self.zero_var = 0
        self.zero_var = self.add_witness_to_circuit_description(0)
        self._add_dummy_constraints()

    # -- variables
    def add_input(self, s: int) -> int:
        self.variables.append(s % Q)
        return len(self.variables) - 1

    def add_witness_to_circuit_description(self, v: int) -> int:
        var = self.add_input(v)
        self.constrain_to_constant(var, v, None)
        return var

    # -- rows
    def _push_row(self, a, b, c, d, q_m, q_l, q_r, q_o, q_4, q_c, pi):
        self.w_l.append(a); self.w_r.append(b); self.w_o.append(c); self.w_4.append(d)
        for wire, var in enumerate((a, b, c, d)):   # perm.add_variables_to_map(a, b, c, d, n)
            self.perm_map.setdefault(var, []).append((self.n, wire))
        vals = dict(q_m=q_m, q_l=q_l, q_r=q_r, q_o=q_o, q_4=q_4, q_c=q_c, q_arith=1,
                    q_range=0, q_logic=0, q_fixed_group_add=0, q_variable_group_add=0)
        for k in SELECTORS:
            self.sel[k].append(vals[k] % Q)
        if pi is not None:
            assert self.n not in self.public_inputs, "duplicate PI position"
            self.public_inputs[self.n] = pi % Q
        self.n += 1

    def poly_gate(self, a, b, c, q_m, q_l, q_r, q_o, q_c, pi):
        self._push_row(a, b, c, self.zero_var, q_m, q_l, q_r, q_o, 0, q_c, pi)
        return (a, b, c)

    def constrain_to_constant(self, a, constant, pi):
        self.poly_gate(a, a, a, 0, 1, 0, 0, -constant, pi)

    def assert_equal(self, a, b):
        self.poly_gate(a, b, self.zero_var, 0, 1, -1, 0, 0, None)

    def big_add(self, q_l_a, q_r_b, q_4_d, q_c, pi):
        q_l, a = q_l_a
        q_r, b = q_r_b
        q_4, d = q_4_d if q_4_d is not None else (0, self.zero_var)
        p = 0 if pi is None else pi
        c_eval = (q_l * self.variables[a] + q_r * self.variables[b] + q_4 * self.variables[d] + q_c + p) % Q
        c = self.add_input(c_eval)
        self._push_row(a, b, c, d, 0, q_l, q_r, -1, q_4, q_c, pi)
        return c

    def add(self, q_l_a, q_r_b, q_c, pi):
        return self.big_add(q_l_a, q_r_b, None, q_c, pi)

    def big_mul(self, q_m, a, b, q_4_d, q_c, pi):
        q_4, d = q_4_d if q_4_d is not None else (0, self.zero_var)
        p = 0 if pi is None else pi
        c_eval = (q_m * self.variables[a] * self.variables[b] + q_4 * self.variables[d] + q_c + p) % Q
        c = self.add_input(c_eval)
        self._push_row(a, b, c, d, q_m, 0, 0, -1, q_4, q_c, pi)
        return c

    def mul(self, q_m, a, b, q_c, pi):
        return self.big_mul(q_m, a, b, None, q_c, pi)

    def mul_gate(self, a, b, c, q_m, q_o, q_c, pi):
        self._push_row(a, b, c, self.zero_var, q_m, 0, 0, q_o, 0, q_c, pi)
        return c

    def boolean_gate(self, a):
        self._push_row(a, a, a, self.zero_var, 1, 0, 0, -1, 0, 0, None)
        return a

    def range_gate(self, witness: int, num_bits: int):
        """StandardComposer::range_gate(witness, num_bits) [dusk-plonk 0.8 src/constraint_system/range.rs, recalled] -- the native
        quad-accumulator range gate the reference recommends for power-of-two bounds (/root/reference/src/range.rs:9-12).
        Accumulators a_i = 4*a_{i-1} + quad_i over the base-4 digits of the witness, most significant first, four per gate in
        the order w_4, w_o, w_r, w_l; q_range = 1 on every used gate but the last; the last accumulator is tied to the
        witness by assert_equal.  Returns the accumulator Variables (the last one replaced by the witness)."""
        assert num_bits % 2 == 0, "number of bits must be even"
        columns = (self.w_4, self.w_o, self.w_r, self.w_l)      # i % 4 = 0 -> fourth, 1 -> output, 2 -> right, 3 -> left
        wire_of = (3, 2, 1, 0)
        base = self.n

        def add_wire(i, variable):
            columns[i % 4].append(variable)
            self.perm_map.setdefault(variable, []).append((base + i // 4, wire_of[i % 4]))

        value = self.variables[witness]
        bits = [(value >> b) & 1 for b in range(256)]            # BitIterator8 over to_bytes(), reversed: LSB first
        num_gates = num_bits >> 3
        if num_bits % 8 != 0:
            num_gates += 1
        num_quads = num_gates * 4
        pad = 1 + (((num_quads << 1) - num_bits) >> 1)
        used_gates = num_gates + 1
        accumulators = []
        accumulator = 0
        for i in range(pad):
            add_wire(i, self.zero_var)
        for i in range(pad, num_quads + 1):
            bit_index = (num_quads - i) << 1
            quad = bits[bit_index] + 2 * bits[bit_index + 1]
            accumulator = (4 * accumulator + quad) % Q
            accumulator_var = self.add_input(accumulator)
            accumulators.append(accumulator_var)
            add_wire(i, accumulator_var)
        for k in SELECTORS:
            self.sel[k].extend([1 if k == "q_range" else 0] * used_gates)
        self.n += used_gates
        self.sel["q_range"][-1] = 0                              # the range equation looks at the next gate's fourth wire
        self.w_l.append(self.zero_var); self.w_r.append(self.zero_var); self.w_o.append(self.zero_var)   # not entered in the permutation map
        last = len(accumulators) - 1
        self.assert_equal(accumulators[last], witness)
        accumulators[last] = witness
        return accumulators

    def _add_dummy_constraints(self):
        var_six = self.add_input(6)
        var_one = self.add_input(1)
        var_seven = self.add_input(7)
        var_min_twenty = self.add_input(-20)
        self._push_row(var_six, var_seven, var_min_twenty, var_one, 1, 2, 3, 4, 1, 4, None)
        self._push_row(var_min_twenty, var_six, var_seven, self.zero_var, 1, 1, 1, 1, 0, 127, None)

    # -- verdict
    def gate_value(self, i: int) -> int:
        a = self.variables[self.w_l[i]]; b = self.variables[self.w_r[i]]
        c = self.variables[self.w_o[i]]; d = self.variables[self.w_4[i]]
        s = self.sel
        pi = self.public_inputs.get(i, 0)
        d_next = self.variables[self.w_4[(i + 1) % self.n]]

        def delta(f):                              # zero iff f is a base-4 digit
            return f * (f - 1) * (f - 2) * (f - 3)
        return (s["q_arith"][i] * (s["q_m"][i] * a * b + s["q_l"][i] * a + s["q_r"][i] * b
                                   + s["q_o"][i] * c + s["q_4"][i] * d + pi + s["q_c"][i])
                + s["q_range"][i] * (delta(c - 4 * d) + delta(b - 4 * c) + delta(a - 4 * b) + delta(d_next - 4 * a))) % Q

    def unsatisfied_rows(self):
        return [i for i in range(self.n) if self.gate_value(i) != 0]

    def construct_dense_pi_vec(self):
        v = [0] * self.n
        for k, x in self.public_inputs.items():
            v[k] = x
        return v

    # -- canonical dump used for golden digests (LE 32-byte canonical scalars, u64 LE indices)
    def digest(self) -> str:
        h = hashlib.sha256()
        h.update(len(self.variables).to_bytes(8, "little"))
        h.update(self.n.to_bytes(8, "little"))
        for v in self.variables:
            h.update(to_bytes(v))
        for i in range(self.n):
            for w in (self.w_l, self.w_r, self.w_o, self.w_4):
                h.update(w[i].to_bytes(8, "little"))
            for k in SELECTORS:
                h.update(to_bytes(self.sel[k][i]))
            h.update(to_bytes(self.public_inputs.get(i, 0)))
        return h.hexdigest()


# --------------------------------------------------------------------------- scalar.rs
def conditionally_select_zero(composer, x, select):
    """scalar.rs:21-27."""
    return composer.mul(1, x, select, 0, None)


def conditionally_select_one(composer, y, selector):
    """scalar.rs:36-59."""
    one = composer.add_witness_to_circuit_description(1)
    selector_y = composer.mul(1, y, selector, 0, None)
    one_min_selector = composer.add((1, one), (-1, selector), 0, None)
    return composer.add((1, selector_y), (1, one_min_selector), 0, None)


def is_non_zero(composer, var, value_assigned):
    """scalar.rs:63-97 (raises NonExistingInverse after 1 var + 1 row were appended)."""
    var_assigned = composer.add_input(value_assigned)
    composer.assert_equal(var, var_assigned)
    inverse = inv_or_none(value_assigned)
    if inverse is None:
        raise NonExistingInverse()
    inv = composer.add_input(inverse)
    one = composer.add_witness_to_circuit_description(1)
    composer.poly_gate(var, inv, one, 1, 0, 0, -1, 0, None)


def maybe_equal(composer, a: AllocatedScalar, b: AllocatedScalar):
    """scalar.rs:105-140."""
    u = composer.add((1, a.var), (-1, b.var), 0, None)
    u_scalar = (a.scalar - b.scalar) % Q
    u_inv = inv_or_none(u_scalar)
    z = composer.add_input(0 if u_inv is None else u_inv)
    y = composer.mul(-1, z, u, 1, None)
    composer.mul_gate(y, u, u, 1, 0, 0, None)
    return y


# --------------------------------------------------------------------------- range.rs
def scalar_decomposition_gadget(composer, num_bits: int, witness: AllocatedScalar):
    """range.rs:119-158."""
    scalar_bits = scalar_to_bits(witness.scalar)
    scalar_bits_var = [composer.add_input(bit) for bit in scalar_bits]
    scalar_bits_var = scalar_bits_var[:num_bits]
    acc = AllocatedScalar(composer.add_witness_to_circuit_description(0), 0)
    for power, bit in enumerate(scalar_bits_var):
        composer.boolean_gate(bit)
        two_pow = pow(2, power, Q)
        acc_var = composer.add((two_pow, bit), (1, acc.var), 0, None)
        acc = AllocatedScalar(acc_var, (acc.scalar + two_pow * scalar_bits[power]) % Q)
    is_equal = maybe_equal(composer, acc, witness)
    return is_equal, scalar_bits_var


def range_proof(composer, value: AllocatedScalar, num_bits: int):
    """range.rs:21-24."""
    is_equal, _ = scalar_decomposition_gadget(composer, num_bits, value)
    return is_equal


def max_bound(composer, max_range: int, witness: AllocatedScalar):
    """range.rs:82-113."""
    max_range = (max_range - 1) % Q
    num_bits_pow_2 = num_bits_closest_power_of_two(max_range)
    b_minus_x_var = composer.add((-1, witness.var), (0, witness.var), max_range, None)
    b_minus_x_scalar = (max_range - witness.scalar) % Q
    return range_proof(composer, AllocatedScalar(b_minus_x_var, b_minus_x_scalar), num_bits_pow_2), num_bits_pow_2


def min_bound(composer, min_range: int, witness: AllocatedScalar, num_bits: int):
    """range.rs:53-76."""
    x_min_a_var = composer.add((1, witness.var), (0, witness.var), -min_range, None)
    x_min_a_scalar = (witness.scalar - min_range) % Q
    return range_proof(composer, AllocatedScalar(x_min_a_var, x_min_a_scalar), num_bits)


def range_check(composer, min_range: int, max_range: int, witness: AllocatedScalar):
    """range.rs:27-43."""
    y1, num_bits = max_bound(composer, max_range, witness)
    y2 = min_bound(composer, min_range, witness, num_bits)
    return composer.mul(1, y1, y2, 0, None)


# ------------------------------------------------------------------------------------------------ evaluation domain
# [DEP] dusk-plonk 0.8 src/fft/domain.rs (EvaluationDomain) and the first step of Prover::prove, reached from
# /root/reference/tests/range_gadgets_tests.rs:90-91.  Independent of oracle/fft.c: plain O(n^2) DFT sums on Python ints.
TWO_ADICITY = 32
GENERATOR = 7
ROOT_OF_UNITY = pow(GENERATOR, (Q - 1) >> TWO_ADICITY, Q)


def domain_log_size(circuit_size: int) -> int:
    return max(circuit_size - 1, 0).bit_length()          # log2(next_power_of_two)


def group_gen(log_n: int) -> int:
    assert log_n <= TWO_ADICITY
    return pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - log_n), Q)


def dft(values, inverse: bool = False):
    """EvaluationDomain::fft (A_k = sum_j a_j w^(jk)) / ifft (w^-1 and a factor n^-1) by the defining sums."""
    n = len(values)
    assert n and n & (n - 1) == 0
    log_n = n.bit_length() - 1
    w = group_gen(log_n)
    if inverse:
        w = pow(w, Q - 2, Q)
    pw = [pow(w, e, Q) for e in range(n)]
    out = [sum(values[j] * pw[(j * k) % n] for j in range(n)) % Q for k in range(n)]
    if inverse:
        n_inv = pow(n, Q - 2, Q)
        out = [x * n_inv % Q for x in out]
    return out


def wire_polynomials(composer: "StandardComposer"):
    """Coefficient vectors of w_l, w_r, w_o, w_4: wire values zero-padded to the domain size, then ifft."""
    size = 1 << domain_log_size(composer.n)
    cols = []
    for wires in (composer.w_l, composer.w_r, composer.w_o, composer.w_4):
        vals = [composer.variables[v] for v in wires] + [0] * (size - composer.n)
        cols.append(dft(vals, inverse=True))
    return cols


# ------------------------------------------------------------------------------------------------ G1 / commitments
# [DEP] dusk-bls12_381 G1Affine/G1Projective and dusk-plonk 0.8 `CommitKey::commit` (src/commitment_scheme/kzg10/key.rs):
#     commitment = msm_variable_base(&self.powers_of_g, &polynomial.coeffs)
# reached from /root/reference/tests/range_gadgets_tests.rs:90-91 through Prover::prove (w_l_poly_commit ... w_4_poly_commit).
# Independent of oracle/g1.c: affine chord-and-tangent formulas on Python ints, one field inversion per addition.
P_FIELD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
G1_B = 4
G1_GENERATOR = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
                0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)


def g1_on_curve(pt) -> bool:
    return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - G1_B) % P_FIELD == 0


def g1_add(a, b):
    """Affine addition; None is the point at infinity."""
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P_FIELD == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P_FIELD) % P_FIELD
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, P_FIELD) % P_FIELD
    x3 = (lam * lam - a[0] - b[0]) % P_FIELD
    return x3, (lam * (a[0] - x3) - a[1]) % P_FIELD


def g1_neg(a):
    return None if a is None else (a[0], (-a[1]) % P_FIELD)


def g1_mul(k: int, pt):
    k %= Q
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, pt)
        pt = g1_add(pt, pt)
        k >>= 1
    return acc


def g1_msm(scalars, points):
    """sum_i scalars[i] * points[i] by the definition (msm_variable_base)."""
    acc = None
    for k, pt in zip(scalars, points):
        acc = g1_add(acc, g1_mul(k, pt))
    return acc


def srs_powers(beta: int, n: int, base=G1_GENERATOR):
    """PublicParameters::setup: powers_of_g[i] = beta^i * g (util::powers_of + slow_multiscalar_mul_single_base)."""
    out, e = [], 1
    for _ in range(n):
        out.append(g1_mul(e, base))
        e = e * beta % Q
    return out


def lagrange_at(beta: int, log_n: int):
    """L_i(beta), i < 2^log_n, for the domain generated by group_gen(log_n), from the defining products
    L_i(X) = prod_{j != i} (X - w^j) / (w^i - w^j)."""
    n = 1 << log_n
    w = group_gen(log_n)
    pts = [pow(w, i, Q) for i in range(n)]
    out = []
    for i in range(n):
        num = den = 1
        for j in range(n):
            if j != i:
                num = num * (beta - pts[j]) % Q
                den = den * (pts[i] - pts[j]) % Q
        out.append(num * pow(den, -1, Q) % Q)
    return out


def srs_lagrange(beta: int, log_n: int, base=G1_GENERATOR):
    return [g1_mul(s, base) for s in lagrange_at(beta, log_n)]
