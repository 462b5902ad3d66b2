"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product package ``plonk_gadgets_b200`` never does.

Scalars cross this boundary as numpy arrays of dtype uint64 and shape (..., 4): the raw little-endian limbs of the
Montgomery form, i.e. exactly the bytes of a ``BlsScalar``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

SEL_NAMES = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_arith", "q_range", "q_logic",
             "q_fixed_group_add", "q_variable_group_add")
Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
ERR_NON_EXISTING_INVERSE = 1
FAITHFUL, FAST = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc).  Idempotent."""
    # make decides by timestamps (sources newer than the .so => rebuild); a prebuilt .so stays usable where make / gcc are missing
    try:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        if force or not os.path.exists(_SO):
            raise
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
        sig = {
            "orc_composer_new": (vp, []), "orc_composer_free": (None, [vp]),
            "orc_n_rows": (u64, [vp]), "orc_n_vars": (u64, [vp]), "orc_wire_ptr": (vp, [vp, i32]),
            "orc_sel_ptr": (vp, [vp, i32]), "orc_dump_variables": (None, [vp, vp]), "orc_value_of_api": (None, [vp, u64, vp]),
            "orc_n_public_inputs": (u64, [vp]), "orc_perm_of": (u64, [vp, u64, vp, u64]),
            "orc_dense_pi": (None, [vp, vp]), "orc_check": (u64, [vp, vp]),
            "orc_add_input_api": (u64, [vp, vp]), "orc_constrain_to_constant_api": (None, [vp, u64, vp, vp]),
            "orc_range_check_api": (u64, [vp, vp, vp, u64]), "orc_max_bound_api": (u64, [vp, vp, u64, vp]),
            "orc_decomposition_api": (u64, [vp, u64, u64]),
            "orc_add_input_batch": (None, [vp, u64, vp, vp]),
            "orc_range_check_batch": (None, [vp, u64, vp, vp, i32, vp, vp]),
            "orc_max_bound_batch": (None, [vp, u64, vp, i32, vp, vp, vp]),
            "orc_maybe_equal_batch": (None, [vp, u64, vp, vp, vp]),
            "orc_is_non_zero_batch": (i32, [vp, u64, vp, vp, vp]),
            "orc_select_zero_batch": (None, [vp, u64, vp, vp, vp]),
            "orc_select_one_batch": (None, [vp, u64, vp, vp, vp]),
            "orc_constrain_to_constant_batch": (None, [vp, u64, vp, vp, vp, i32]),
            "orc_range_gate_batch": (None, [vp, u64, vp, u64]),
            "orc_poly_gate_batch": (i32, [vp, u64, vp, vp, vp, vp, vp]),
            "orc_set_mode": (None, [i32]),
            "orc_fr_from_u64": (None, [u64, vp]), "orc_fr_mul": (None, [vp, vp, vp]), "orc_fr_add": (None, [vp, vp, vp]),
            "orc_fr_sub": (None, [vp, vp, vp]), "orc_fr_neg": (None, [vp, vp]), "orc_fr_invert": (i32, [vp, vp]),
            "orc_fr_pow": (None, [vp, vp, vp]), "orc_fr_pow_of_2": (None, [u64, vp]), "orc_fr_reduce": (None, [vp, vp]),
            "orc_fr_to_bytes": (None, [vp, vp]), "orc_fr_from_bytes": (i32, [vp, vp]), "orc_fr_from_bytes_wide": (None, [vp, vp]),
            "orc_fr_from_bytes_many": (i32, [u64, vp, vp]), "orc_fr_to_bytes_many": (None, [u64, vp, vp]),
            "orc_fr_from_bytes_wide_many": (None, [u64, vp, vp]),
            "orc_bits_count_api": (u64, [vp]), "orc_num_bits_api": (u64, [vp]),
            "orc_fft": (i32, [vp, C.c_uint, i32]), "orc_domain_group_gen": (None, [C.c_uint, vp]),
            "orc_domain_log_size": (C.c_uint, [u64]), "orc_wire_polynomials": (C.c_uint, [vp, vp]),
            "orc_fp_from_raw": (None, [vp, vp]), "orc_fp_to_raw": (None, [vp, vp]), "orc_fp_mul": (None, [vp, vp, vp]),
            "orc_fp_add": (None, [vp, vp, vp]), "orc_fp_sub": (None, [vp, vp, vp]), "orc_fp_neg": (None, [vp, vp]), "orc_fp_inv": (None, [vp, vp]),
            "orc_g1_generator": (None, [vp]), "orc_g1_on_curve": (i32, [vp]), "orc_g1_add": (None, [vp, vp, vp]),
            "orc_g1_mul": (None, [vp, vp, vp]), "orc_g1_msm": (None, [u64, vp, vp, vp]), "orc_srs_powers": (None, [vp, vp, u64, vp]),
            "orc_bench_range": (dbl, [i32, u64, vp, vp, vp, i32, i32, i32, u64, vp, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _fr_arr(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


# ------------------------------------------------------------------ conversions
def from_ints(vals) -> np.ndarray:
    """Canonical Python ints (reduced mod q) -> (n,4) uint64 Montgomery limbs, through the oracle's from_bytes."""
    vals = [int(v) % Q for v in vals]
    raw = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), dtype=np.uint8)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    if len(vals):
        ok = lib().orc_fr_from_bytes_many(len(vals), _p(raw), _p(out))
        assert ok
    return out


def to_ints(a) -> list:
    """(n,4) Montgomery limbs -> canonical Python ints, through the oracle's to_bytes."""
    a = _fr_arr(a).reshape(-1, 4)
    out = np.empty((a.shape[0], 32), dtype=np.uint8)
    if a.shape[0]:
        lib().orc_fr_to_bytes_many(a.shape[0], _p(a), _p(out))
    return [int.from_bytes(out[i].tobytes(), "little") for i in range(a.shape[0])]


def from_bytes_wide(raw: np.ndarray) -> np.ndarray:
    """(n,64) uint8 -> (n,4) Montgomery limbs (BlsScalar::from_bytes_wide: uniform mod q)."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, 64)
    out = np.empty((raw.shape[0], 4), dtype=np.uint64)
    if raw.shape[0]:
        lib().orc_fr_from_bytes_wide_many(raw.shape[0], _p(raw), _p(out))
    return out


# ------------------------------------------------------------------ composer
class Composer:
    """The oracle's StandardComposer (fresh: 3 rows / 5 variables)."""

    def __init__(self):
        self._c = C.c_void_p(lib().orc_composer_new())

    def __del__(self):
        if getattr(self, "_c", None):
            lib().orc_composer_free(self._c)
            self._c = None

    # state
    @property
    def n(self) -> int:
        return lib().orc_n_rows(self._c)

    @property
    def n_vars(self) -> int:
        return lib().orc_n_vars(self._c)

    def variables(self) -> np.ndarray:
        out = np.empty((self.n_vars, 4), dtype=np.uint64)
        lib().orc_dump_variables(self._c, _p(out))
        return out

    def wires(self) -> np.ndarray:
        """(4, n) uint64 Variable indices: w_l, w_r, w_o, w_4."""
        n = self.n
        out = np.empty((4, n), dtype=np.uint64)
        for w in range(4):
            ptr = lib().orc_wire_ptr(self._c, w)
            out[w] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,))
        return out

    def selectors(self) -> np.ndarray:
        """(11, n, 4) uint64 in SEL_NAMES order."""
        n = self.n
        out = np.empty((len(SEL_NAMES), n, 4), dtype=np.uint64)
        for s in range(len(SEL_NAMES)):
            ptr = lib().orc_sel_ptr(self._c, s)
            out[s] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n, 4))
        return out

    def dense_pi(self) -> np.ndarray:
        out = np.empty((self.n, 4), dtype=np.uint64)
        lib().orc_dense_pi(self._c, _p(out))
        return out

    def check(self):
        """(number of unsatisfied rows, first unsatisfied row or None)."""
        fb = C.c_uint64(0)
        bad = lib().orc_check(self._c, C.byref(fb))
        return bad, (None if fb.value == 2 ** 64 - 1 else fb.value)

    def perm_of(self, var: int):
        buf = np.empty(64, dtype=np.uint64)
        cnt = lib().orc_perm_of(self._c, var, _p(buf), 64)
        if cnt > 64:
            buf = np.empty(cnt, dtype=np.uint64)
            lib().orc_perm_of(self._c, var, _p(buf), cnt)
        return [(int(x) // 4, int(x) % 4) for x in buf[:cnt]]

    # batch programs: return arrays of Variable indices
    def add_input_batch(self, vals) -> np.ndarray:
        vals = _fr_arr(vals).reshape(-1, 4)
        out = np.empty(vals.shape[0], dtype=np.uint64)
        lib().orc_add_input_batch(self._c, vals.shape[0], _p(vals), _p(out))
        return out

    @staticmethod
    def _bounds(b, n):
        b = _fr_arr(b).reshape(-1, 4)
        uniform = b.shape[0] == 1
        assert uniform or b.shape[0] == n
        return b, int(uniform)

    def range_check_batch(self, mn, mx, wit) -> np.ndarray:
        wit = np.ascontiguousarray(wit, dtype=np.uint64)
        mn, u1 = self._bounds(mn, len(wit)); mx, u2 = self._bounds(mx, len(wit))
        assert u1 == u2
        out = np.empty(len(wit), dtype=np.uint64)
        lib().orc_range_check_batch(self._c, len(wit), _p(mn), _p(mx), u1, _p(wit), _p(out))
        return out

    def max_bound_batch(self, mx, wit):
        wit = np.ascontiguousarray(wit, dtype=np.uint64)
        mx, u = self._bounds(mx, len(wit))
        out = np.empty(len(wit), dtype=np.uint64)
        nb = np.empty(len(wit), dtype=np.uint64)
        lib().orc_max_bound_batch(self._c, len(wit), _p(mx), u, _p(wit), _p(out), _p(nb))
        return out, nb

    def maybe_equal_batch(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.empty(len(a), dtype=np.uint64)
        lib().orc_maybe_equal_batch(self._c, len(a), _p(a), _p(b), _p(out))
        return out

    def is_non_zero_batch(self, vars_, assigned):
        """Returns (error code, number of completed instances) -- stops at the first zero like `?`."""
        vars_ = np.ascontiguousarray(vars_, dtype=np.uint64)
        assigned = _fr_arr(assigned).reshape(-1, 4)
        done = C.c_uint64(0)
        e = lib().orc_is_non_zero_batch(self._c, len(vars_), _p(vars_), _p(assigned), C.byref(done))
        return e, done.value

    def select_zero_batch(self, x, s) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.uint64); s = np.ascontiguousarray(s, dtype=np.uint64)
        out = np.empty(len(x), dtype=np.uint64)
        lib().orc_select_zero_batch(self._c, len(x), _p(x), _p(s), _p(out))
        return out

    def select_one_batch(self, y, s) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.uint64); s = np.ascontiguousarray(s, dtype=np.uint64)
        out = np.empty(len(y), dtype=np.uint64)
        lib().orc_select_one_batch(self._c, len(y), _p(y), _p(s), _p(out))
        return out

    def range_gate_batch(self, wit, num_bits: int):
        wit = np.ascontiguousarray(wit, dtype=np.uint64)
        lib().orc_range_gate_batch(self._c, len(wit), _p(wit), int(num_bits))

    def poly_gate_batch(self, a, b, o, sel6, pi):
        """n x poly_gate(a, b, o, q_m, q_l, q_r, q_o, q_c, pi): sel6 (6, n, 4) in q_m q_l q_r q_o q_4 q_c order (q_4 must be zero)."""
        a, b, o = (np.ascontiguousarray(x, dtype=np.uint64) for x in (a, b, o))
        sel6 = _fr_arr(sel6).reshape(6, len(a), 4); pi = _fr_arr(pi).reshape(len(a), 4)
        if lib().orc_poly_gate_batch(self._c, len(a), _p(a), _p(b), _p(o), _p(sel6), _p(pi)) != 0:
            raise ValueError("poly_gate cannot express a row with q_4 != 0")

    def constrain_to_constant_batch(self, vars_, k, pi=None):
        vars_ = np.ascontiguousarray(vars_, dtype=np.uint64)
        k, u = self._bounds(k, len(vars_))
        if pi is not None:
            pi, u2 = self._bounds(pi, len(vars_))
            if u != u2:   # one of them is per-instance: broadcast the other
                k = np.ascontiguousarray(np.broadcast_to(k, (len(vars_), 4)))
                pi = np.ascontiguousarray(np.broadcast_to(pi, (len(vars_), 4)))
                u = 0
        lib().orc_constrain_to_constant_batch(self._c, len(vars_), _p(vars_), _p(k), _p(pi) if pi is not None else None, u)

    def wire_polynomials(self) -> np.ndarray:
        """(4, 2^k, 4) uint64: coefficients of w_l, w_r, w_o, w_4 over the padded evaluation domain (oracle/fft.c)."""
        L = lib()
        size = 1 << L.orc_domain_log_size(self.n)
        out = np.zeros((4, size, 4), dtype=np.uint64)
        L.orc_wire_polynomials(self._c, _p(out))
        return out

    def decomposition(self, num_bits: int, var: int) -> int:
        return lib().orc_decomposition_api(self._c, num_bits, var)


def fft(a, inverse: bool = False) -> np.ndarray:
    """EvaluationDomain::fft / ifft of a vector of 2^k Montgomery scalars ((n, 4) uint64); returns a new array."""
    a = np.ascontiguousarray(_fr_arr(a)).copy()
    n = a.shape[0]
    assert n and n & (n - 1) == 0
    rc = lib().orc_fft(_p(a), n.bit_length() - 1, 1 if inverse else 0)
    assert rc == 0
    return a


def domain_group_gen(log_n: int) -> int:
    out = np.zeros((1, 4), dtype=np.uint64)
    lib().orc_domain_group_gen(log_n, _p(out))
    return to_ints(out)[0]


# ---- G1 (oracle/g1.c).  A point is 13 uint64: 6 limbs of x, 6 limbs of y (Montgomery form, like the crate's Fp), infinity flag.
P_FIELD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_FIELD = (1 << 384) % P_FIELD


def g1_from_ints(points) -> np.ndarray:
    """[(x, y) | None, ...] canonical integers -> (n, 13) uint64 oracle points."""
    out = np.zeros((len(points), 13), dtype=np.uint64)
    for i, pt in enumerate(points):
        if pt is None:
            out[i, 12] = 1
            continue
        for c, v in enumerate(pt):
            m = v * R_FIELD % P_FIELD
            for k in range(6):
                out[i, 6 * c + k] = (m >> (64 * k)) & (2 ** 64 - 1)
    return out


def g1_to_ints(a: np.ndarray) -> list:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 13)
    r_inv = pow(R_FIELD, -1, P_FIELD)
    out = []
    for row in a:
        if int(row[12]) & 0xffffffff:
            out.append(None)
            continue
        x = sum(int(row[k]) << (64 * k) for k in range(6)) * r_inv % P_FIELD
        y = sum(int(row[6 + k]) << (64 * k) for k in range(6)) * r_inv % P_FIELD
        out.append((x, y))
    return out


def g1_generator() -> np.ndarray:
    out = np.zeros((1, 13), dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 13); b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 13)
    out = np.zeros_like(a)
    for i in range(a.shape[0]):
        lib().orc_g1_add(_p(a[i:i + 1]), _p(b[i:i + 1]), _p(out[i:i + 1]))
    return out


def g1_mul(a: np.ndarray, k) -> np.ndarray:
    """k: (n, 4) Montgomery scalars."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 13); k = np.ascontiguousarray(_fr_arr(k))
    out = np.zeros_like(a)
    for i in range(a.shape[0]):
        lib().orc_g1_mul(_p(a[i:i + 1]), _p(k[i:i + 1]), _p(out[i:i + 1]))
    return out


def g1_msm(points: np.ndarray, scalars) -> np.ndarray:
    """msm_variable_base(points, scalars) -> one point."""
    points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 13); scalars = np.ascontiguousarray(_fr_arr(scalars))
    assert points.shape[0] == scalars.shape[0]
    out = np.zeros((1, 13), dtype=np.uint64)
    lib().orc_g1_msm(points.shape[0], _p(points), _p(scalars), _p(out))
    return out


def srs_powers(beta, n: int, base: np.ndarray | None = None) -> np.ndarray:
    """powers_of_g[i] = beta^i * g (PublicParameters::setup); beta: one Montgomery scalar."""
    base = g1_generator() if base is None else np.ascontiguousarray(base, dtype=np.uint64).reshape(1, 13)
    beta = np.ascontiguousarray(_fr_arr(beta))
    out = np.zeros((n, 13), dtype=np.uint64)
    lib().orc_srs_powers(_p(beta), _p(base), n, _p(out))
    return out


def bench_range(gadget: int, wit, mn, mx, threads: int, mode: int = FAITHFUL, chunk: int = 64, want_results=False):
    """Timed CPU baseline (witness generation through the composer + gate check).  Returns a dict."""
    wit = _fr_arr(wit).reshape(-1, 4)
    mn = _fr_arr(mn).reshape(-1, 4); mx = _fr_arr(mx).reshape(-1, 4)
    uniform = int(mx.shape[0] == 1)
    res = np.empty((wit.shape[0], 4), dtype=np.uint64) if want_results else None
    rows, unsat, ones = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    sec = lib().orc_bench_range(gadget, wit.shape[0], _p(wit), _p(mn), _p(mx), uniform, threads, mode, chunk,
                                _p(res) if res is not None else None, C.byref(rows), C.byref(unsat), C.byref(ones))
    return dict(seconds=sec, rows=rows.value, unsat=unsat.value, ones=ones.value, results=res)
