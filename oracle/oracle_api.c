/*
 * oracle/oracle_api.c -- flat C entry points (for ctypes) over the CPU oracle.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
 * product library.
 *
 * The *_batch functions run the sequential reference program a batched engine call is defined to be equal to:
 *     for i in 0..n { gadget(composer, ..., operand_i) }
 * with the reference's own early-exit for `is_non_zero(..)?` (/root/reference/src/scalar.rs:79).
 */
#include <string.h>
#include "gadgets.h"
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

/* ---- Fr helpers (L0 tests) ---- */
void orc_fr_from_u64(uint64_t v, fr_t *out) { *out = fr_from_u64(v); }
void orc_fr_mul(const fr_t *a, const fr_t *b, fr_t *out) { *out = fr_mul(a, b); }
void orc_fr_add(const fr_t *a, const fr_t *b, fr_t *out) { *out = fr_add(a, b); }
void orc_fr_sub(const fr_t *a, const fr_t *b, fr_t *out) { *out = fr_sub(a, b); }
void orc_fr_neg(const fr_t *a, fr_t *out) { *out = fr_neg(a); }
int orc_fr_invert(const fr_t *a, fr_t *out) { return fr_invert(a, out); }
void orc_fr_pow(const fr_t *a, const uint64_t by[4], fr_t *out) { *out = fr_pow(a, by); }
void orc_fr_pow_of_2(uint64_t by, fr_t *out) { *out = fr_pow_of_2(by); }
void orc_fr_reduce(const fr_t *a, fr_t *out) { *out = fr_reduce(a); }
void orc_fr_to_bytes(const fr_t *a, uint8_t out[32]) { fr_to_bytes(a, out); }
int orc_fr_from_bytes(const uint8_t in[32], fr_t *out) { return fr_from_bytes(in, out); }
void orc_fr_from_bytes_wide(const uint8_t in[64], fr_t *out) { *out = fr_from_bytes_wide(in); }
/* n canonical 32-byte LE values -> n Montgomery scalars (inputs must be < q) */
int orc_fr_from_bytes_many(uint64_t n, const uint8_t *in, fr_t *out) {
    for (uint64_t i = 0; i < n; i++) if (!fr_from_bytes(in + 32 * i, &out[i])) return 0;
    return 1;
}
void orc_fr_to_bytes_many(uint64_t n, const fr_t *in, uint8_t *out) { for (uint64_t i = 0; i < n; i++) fr_to_bytes(&in[i], out + 32 * i); }
void orc_fr_from_bytes_wide_many(uint64_t n, const uint8_t *in, fr_t *out) { for (uint64_t i = 0; i < n; i++) out[i] = fr_from_bytes_wide(in + 64 * i); }
uint64_t orc_bits_count_api(const fr_t *a) { return orc_bits_count(*a); }
uint64_t orc_num_bits_api(const fr_t *a) { return orc_num_bits_closest_power_of_two(*a); }

/* ---- accessors ---- */
uint64_t orc_n_rows(const orc_composer *c) { return c->n; }
uint64_t orc_n_vars(const orc_composer *c) { return c->n_vars; }
const uint64_t *orc_wire_ptr(const orc_composer *c, int w) { return c->w[w].p; }
const fr_t *orc_sel_ptr(const orc_composer *c, int s) { return c->sel[s].p; }
void orc_dump_variables(const orc_composer *c, fr_t *out) { for (uint64_t v = 0; v < c->n_vars; v++) out[v] = orc_value_of(c, v); }
void orc_value_of_api(const orc_composer *c, uint64_t var, fr_t *out) { *out = orc_value_of(c, var); }
uint64_t orc_n_public_inputs(const orc_composer *c) { return c->pi_pos.len; }
/* perm.variable_map of one variable: fills up to cap entries (row*4+wire, oldest first), returns the count */
uint64_t orc_perm_of(const orc_composer *c, uint64_t var, uint64_t *out, uint64_t cap) {
    uint64_t cnt = 0;
    for (uint64_t e = c->perm_head.p[var]; e; e = c->perm_next.p[e - 1]) cnt++;
    uint64_t k = cnt;
    for (uint64_t e = c->perm_head.p[var]; e; e = c->perm_next.p[e - 1]) { k--; if (k < cap) out[k] = c->perm_data.p[e - 1]; }
    return cnt;
}

/* ---- single-instance wrappers (pointer arguments only) ---- */
uint64_t orc_add_input_api(orc_composer *c, const fr_t *v) { return orc_add_input(c, v); }
void orc_constrain_to_constant_api(orc_composer *c, uint64_t a, const fr_t *k, const fr_t *pi) { orc_constrain_to_constant(c, a, k, pi); }
uint64_t orc_range_check_api(orc_composer *c, const fr_t *mn, const fr_t *mx, uint64_t var) {
    orc_allocated_scalar w; w.var = var; w.scalar = orc_value_of(c, var); return orc_range_check(c, *mn, *mx, w);
}
uint64_t orc_max_bound_api(orc_composer *c, const fr_t *mx, uint64_t var, uint64_t *k) {
    orc_allocated_scalar w; w.var = var; w.scalar = orc_value_of(c, var); return orc_max_bound(c, *mx, w, k);
}
uint64_t orc_decomposition_api(orc_composer *c, uint64_t num_bits, uint64_t var) {
    orc_allocated_scalar w; w.var = var; w.scalar = orc_value_of(c, var); return orc_scalar_decomposition_gadget(c, num_bits, w);
}

/* ---- batch programs ---- */
void orc_add_input_batch(orc_composer *c, uint64_t n, const fr_t *vals, uint64_t *out_vars) {
    for (uint64_t i = 0; i < n; i++) { uint64_t v = orc_add_input(c, &vals[i]); if (out_vars) out_vars[i] = v; }
}
/* uniform != 0: min[0]/max[0] apply to every instance.  The AllocatedScalar's scalar is the variable's own value. */
void orc_range_check_batch(orc_composer *c, uint64_t n, const fr_t *mn, const fr_t *mx, int uniform, const uint64_t *wit, uint64_t *out_vars) {
    for (uint64_t i = 0; i < n; i++) {
        orc_allocated_scalar w; w.var = wit[i]; w.scalar = orc_value_of(c, wit[i]);
        uint64_t y = orc_range_check(c, mn[uniform ? 0 : i], mx[uniform ? 0 : i], w);
        if (out_vars) out_vars[i] = y;
    }
}
void orc_max_bound_batch(orc_composer *c, uint64_t n, const fr_t *mx, int uniform, const uint64_t *wit, uint64_t *out_vars, uint64_t *num_bits) {
    for (uint64_t i = 0; i < n; i++) {
        orc_allocated_scalar w; w.var = wit[i]; w.scalar = orc_value_of(c, wit[i]);
        uint64_t k, y = orc_max_bound(c, mx[uniform ? 0 : i], w, &k);
        if (out_vars) out_vars[i] = y;
        if (num_bits) num_bits[i] = k;
    }
}
void orc_maybe_equal_batch(orc_composer *c, uint64_t n, const uint64_t *a, const uint64_t *b, uint64_t *out_vars) {
    for (uint64_t i = 0; i < n; i++) {
        orc_allocated_scalar x, y; x.var = a[i]; x.scalar = orc_value_of(c, a[i]); y.var = b[i]; y.scalar = orc_value_of(c, b[i]);
        uint64_t r = orc_maybe_equal(c, x, y); if (out_vars) out_vars[i] = r;
    }
}
/* `for i { is_non_zero(composer, var_i, assigned_i)?; }`: stops at the first error; *n_done = completed instances */
int orc_is_non_zero_batch(orc_composer *c, uint64_t n, const uint64_t *vars, const fr_t *assigned, uint64_t *n_done) {
    for (uint64_t i = 0; i < n; i++) {
        int e = orc_is_non_zero(c, vars[i], assigned[i]);
        if (e) { if (n_done) *n_done = i; return e; }
    }
    if (n_done) *n_done = n;
    return ORC_OK;
}
void orc_select_zero_batch(orc_composer *c, uint64_t n, const uint64_t *x, const uint64_t *s, uint64_t *out_vars) {
    for (uint64_t i = 0; i < n; i++) { uint64_t r = orc_conditionally_select_zero(c, x[i], s[i]); if (out_vars) out_vars[i] = r; }
}
void orc_select_one_batch(orc_composer *c, uint64_t n, const uint64_t *y, const uint64_t *s, uint64_t *out_vars) {
    for (uint64_t i = 0; i < n; i++) { uint64_t r = orc_conditionally_select_one(c, y[i], s[i]); if (out_vars) out_vars[i] = r; }
}
/* `for i { composer.range_gate(witness_i, num_bits); }` */
void orc_range_gate_batch(orc_composer *c, uint64_t n, const uint64_t *wit, uint64_t num_bits) {
    for (uint64_t i = 0; i < n; i++) orc_range_gate(c, wit[i], num_bits);
}
/* pi == NULL: no public input; else pi[i] (or pi[0] when uniform) is attached to row i's PI slot */
void orc_constrain_to_constant_batch(orc_composer *c, uint64_t n, const uint64_t *vars, const fr_t *k, const fr_t *pi, int uniform) {
    for (uint64_t i = 0; i < n; i++) orc_constrain_to_constant(c, vars[i], &k[uniform ? 0 : i], pi ? &pi[uniform ? 0 : i] : NULL);
}

/* replay of exported rows (tests of the import adapter's input): n x poly_gate(a, b, o, q_m, q_l, q_r, q_o, q_c, pi) -- the public
 * StandardComposer method a replaying host uses (call site in the reference: /root/reference/src/scalar.rs:84); sel6 is column-major
 * q_m q_l q_r q_o q_4 q_c (q_4 must be 0: poly_gate has no fourth wire); pi[i] == 0 is replayed as None */
int orc_poly_gate_batch(orc_composer *c, uint64_t n, const uint64_t *a, const uint64_t *b, const uint64_t *o, const fr_t *sel6, const fr_t *pi) {
    const fr_t zero = fr_zero();
    for (uint64_t i = 0; i < n; i++) {
        if (memcmp(&sel6[4 * n + i], &zero, sizeof(fr_t)) != 0) return -1;
        const int has_pi = memcmp(&pi[i], &zero, sizeof(fr_t)) != 0;
        orc_poly_gate(c, a[i], b[i], o[i], &sel6[i], &sel6[n + i], &sel6[2 * n + i], &sel6[3 * n + i], &sel6[5 * n + i], has_pi ? &pi[i] : NULL);
    }
    return 0;
}

/* ---- timed CPU baseline: the reference path (witness generation through the composer + gate check) -------- */
typedef struct {
    uint64_t lo, hi; const fr_t *wit, *mn, *mx; int uniform, gadget; uint64_t chunk;
    uint64_t rows, unsat, ones; fr_t *results;
} bench_job;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

static void *bench_worker(void *arg) {
    bench_job *j = (bench_job *)arg; fr_t one = fr_one();
    for (uint64_t base = j->lo; base < j->hi; base += j->chunk) {
        uint64_t end = base + j->chunk < j->hi ? base + j->chunk : j->hi;
        orc_composer *c = orc_composer_new();
        uint64_t rows0 = c->n;
        for (uint64_t i = base; i < end; i++) {
            orc_allocated_scalar w = orc_allocate(c, &j->wit[i]);
            uint64_t y;
            if (j->gadget == 0) y = orc_range_check(c, j->mn[j->uniform ? 0 : i], j->mx[j->uniform ? 0 : i], w);
            else y = orc_max_bound(c, j->mx[j->uniform ? 0 : i], w, NULL);
            fr_t yv = orc_value_of(c, y);
            if (j->results) j->results[i] = yv;
            if (fr_eq(&yv, &one)) j->ones++;
        }
        uint64_t fb; j->unsat += orc_check(c, &fb);
        j->rows += c->n - rows0;
        orc_composer_free(c);
    }
    return NULL;
}

/* gadget: 0 = range_check, 1 = max_bound.  One composer per thread per `chunk` instances (allocate + gadget for each,
 * then the gate check over everything appended).  Returns wall seconds; *rows = gadget rows generated and checked. */
double orc_bench_range(int gadget, uint64_t n, const fr_t *wit, const fr_t *mn, const fr_t *mx, int uniform, int threads, int mode,
                       uint64_t chunk, fr_t *results, uint64_t *rows, uint64_t *unsat, uint64_t *ones) {
    if (threads < 1) threads = 1;
    if (chunk < 1) chunk = 64;
    orc_set_mode(mode);
    bench_job *jobs = (bench_job *)calloc((size_t)threads, sizeof(bench_job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    double t0 = now_s();
    for (int t = 0; t < threads; t++) {
        jobs[t].lo = n * (uint64_t)t / (uint64_t)threads; jobs[t].hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
        jobs[t].wit = wit; jobs[t].mn = mn; jobs[t].mx = mx; jobs[t].uniform = uniform; jobs[t].gadget = gadget;
        jobs[t].chunk = chunk; jobs[t].results = results;
        pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
    }
    uint64_t r = 0, u = 0, o = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); r += jobs[t].rows; u += jobs[t].unsat; o += jobs[t].ones; }
    double t1 = now_s();
    if (rows) *rows = r;
    if (unsat) *unsat = u;
    if (ones) *ones = o;
    free(jobs); free(th);
    return t1 - t0;
}
