/* oracle/composer.c -- see composer.h.  TEST INFRASTRUCTURE ONLY. */
#include "composer.h"
#include <stdlib.h>
#include <stdio.h>

static void die(const char *m) { fprintf(stderr, "oracle: %s\n", m); abort(); }

static void frvec_push(orc_frvec *v, const fr_t *x) {
    if (v->len == v->cap) { v->cap = v->cap ? v->cap * 2 : 16; v->p = (fr_t *)realloc(v->p, v->cap * sizeof(fr_t)); if (!v->p) die("oom"); }
    v->p[v->len++] = *x;
}
static void u64vec_push(orc_u64vec *v, uint64_t x) {
    if (v->len == v->cap) { v->cap = v->cap ? v->cap * 2 : 16; v->p = (uint64_t *)realloc(v->p, v->cap * sizeof(uint64_t)); if (!v->p) die("oom"); }
    v->p[v->len++] = x;
}

/* ---- HashMap<Variable, BlsScalar> ---------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL; x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL; return x ^ (x >> 31);
}
#define EMPTY_KEY (~(uint64_t)0)
static void map_grow(orc_composer *c) {
    size_t ncap = c->map_cap ? c->map_cap * 2 : 64;
    uint64_t *nk = (uint64_t *)malloc(ncap * sizeof(uint64_t)); fr_t *nv = (fr_t *)malloc(ncap * sizeof(fr_t));
    if (!nk || !nv) die("oom");
    for (size_t i = 0; i < ncap; i++) nk[i] = EMPTY_KEY;
    for (size_t i = 0; i < c->map_cap; i++) if (c->map_key[i] != EMPTY_KEY) {
        size_t h = mix64(c->map_key[i]) & (ncap - 1);
        while (nk[h] != EMPTY_KEY) h = (h + 1) & (ncap - 1);
        nk[h] = c->map_key[i]; nv[h] = c->map_val[i];
    }
    free(c->map_key); free(c->map_val); c->map_key = nk; c->map_val = nv; c->map_cap = ncap;
}
static void map_insert(orc_composer *c, uint64_t key, const fr_t *v) {
    if ((c->map_len + 1) * 8 > c->map_cap * 7) map_grow(c);
    size_t h = mix64(key) & (c->map_cap - 1);
    while (c->map_key[h] != EMPTY_KEY && c->map_key[h] != key) h = (h + 1) & (c->map_cap - 1);
    if (c->map_key[h] == EMPTY_KEY) c->map_len++;
    c->map_key[h] = key; c->map_val[h] = *v;
}
fr_t orc_value_of(const orc_composer *c, uint64_t var) {
    size_t h = mix64(var) & (c->map_cap - 1);
    while (c->map_key[h] != var) { if (c->map_key[h] == EMPTY_KEY) die("unknown Variable"); h = (h + 1) & (c->map_cap - 1); }
    return c->map_val[h];
}

/* ---- rows ------------------------------------------------------------------------------------------ */
static void perm_add(orc_composer *c, uint64_t var, uint64_t row, unsigned wire) {
    u64vec_push(&c->perm_data, row * 4 + wire);
    u64vec_push(&c->perm_next, c->perm_head.p[var]);
    c->perm_head.p[var] = c->perm_data.len;   /* index + 1 */
}
static void push_row(orc_composer *c, uint64_t a, uint64_t b, uint64_t o, uint64_t d, const fr_t *q_m, const fr_t *q_l,
                     const fr_t *q_r, const fr_t *q_o, const fr_t *q_4, const fr_t *q_c, const fr_t *pi) {
    fr_t zero = fr_zero(), one = fr_one();
    u64vec_push(&c->w[0], a); u64vec_push(&c->w[1], b); u64vec_push(&c->w[2], o); u64vec_push(&c->w[3], d);
    frvec_push(&c->sel[ORC_QM], q_m); frvec_push(&c->sel[ORC_QL], q_l); frvec_push(&c->sel[ORC_QR], q_r);
    frvec_push(&c->sel[ORC_QO], q_o); frvec_push(&c->sel[ORC_Q4], q_4); frvec_push(&c->sel[ORC_QC], q_c);
    frvec_push(&c->sel[ORC_QARITH], &one); frvec_push(&c->sel[ORC_QRANGE], &zero); frvec_push(&c->sel[ORC_QLOGIC], &zero);
    frvec_push(&c->sel[ORC_QFIXED], &zero); frvec_push(&c->sel[ORC_QVAR], &zero);
    if (pi) {
        if (c->pi_pos.len && c->pi_pos.p[c->pi_pos.len - 1] == c->n) die("duplicate PI position");   /* the crate assert!s */
        u64vec_push(&c->pi_pos, c->n); frvec_push(&c->pi_val, pi);
    }
    perm_add(c, a, c->n, 0); perm_add(c, b, c->n, 1); perm_add(c, o, c->n, 2); perm_add(c, d, c->n, 3);
    c->n++;
}

uint64_t orc_add_input(orc_composer *c, const fr_t *s) {
    uint64_t var = c->n_vars++;
    map_insert(c, var, s);
    u64vec_push(&c->perm_head, 0);
    return var;
}
void orc_poly_gate(orc_composer *c, uint64_t a, uint64_t b, uint64_t o, const fr_t *q_m, const fr_t *q_l,
                   const fr_t *q_r, const fr_t *q_o, const fr_t *q_c, const fr_t *pi) {
    fr_t zero = fr_zero();
    push_row(c, a, b, o, c->zero_var, q_m, q_l, q_r, q_o, &zero, q_c, pi);
}
void orc_constrain_to_constant(orc_composer *c, uint64_t a, const fr_t *constant, const fr_t *pi) {
    fr_t zero = fr_zero(), one = fr_one(), nc = fr_neg(constant);
    orc_poly_gate(c, a, a, a, &zero, &one, &zero, &zero, &nc, pi);
}
uint64_t orc_add_witness_to_circuit_description(orc_composer *c, const fr_t *v) {
    uint64_t var = orc_add_input(c, v);
    orc_constrain_to_constant(c, var, v, NULL);
    return var;
}
void orc_assert_equal(orc_composer *c, uint64_t a, uint64_t b) {
    fr_t zero = fr_zero(), one = fr_one(), m1 = fr_neg(&one);
    orc_poly_gate(c, a, b, c->zero_var, &zero, &one, &m1, &zero, &zero, NULL);
}
/* add(q_l_a, q_r_b, q_c, pi) = big_add(.., None, ..): c = q_l*a + q_r*b + q_4*d + q_c + pi with d = (0, zero_var) */
uint64_t orc_add(orc_composer *c, const fr_t *q_l, uint64_t a, const fr_t *q_r, uint64_t b, const fr_t *q_c, const fr_t *pi) {
    fr_t zero = fr_zero(), one = fr_one(), q_o = fr_neg(&one), q_4 = zero;
    uint64_t d = c->zero_var;
    fr_t av = orc_value_of(c, a), bv = orc_value_of(c, b), dv = orc_value_of(c, d);
    fr_t t0 = fr_mul(q_l, &av), t1 = fr_mul(q_r, &bv), t2 = fr_mul(&q_4, &dv);
    fr_t s = fr_add(&t0, &t1); s = fr_add(&s, &t2); s = fr_add(&s, q_c);
    fr_t p = pi ? *pi : zero; s = fr_add(&s, &p);
    uint64_t o = orc_add_input(c, &s);
    push_row(c, a, b, o, d, &zero, q_l, q_r, &q_o, &q_4, q_c, pi);
    return o;
}
/* mul(q_m, a, b, q_c, pi) = big_mul(.., None, ..): c = q_m*a*b + q_4*d + q_c + pi */
uint64_t orc_mul(orc_composer *c, const fr_t *q_m, uint64_t a, uint64_t b, const fr_t *q_c, const fr_t *pi) {
    fr_t zero = fr_zero(), one = fr_one(), q_o = fr_neg(&one), q_4 = zero;
    uint64_t d = c->zero_var;
    fr_t av = orc_value_of(c, a), bv = orc_value_of(c, b), dv = orc_value_of(c, d);
    fr_t t0 = fr_mul(&av, &bv); t0 = fr_mul(q_m, &t0);
    fr_t t2 = fr_mul(&q_4, &dv);
    fr_t s = fr_add(&t0, &t2); s = fr_add(&s, q_c);
    fr_t p = pi ? *pi : zero; s = fr_add(&s, &p);
    uint64_t o = orc_add_input(c, &s);
    push_row(c, a, b, o, d, q_m, &zero, &zero, &q_o, &q_4, q_c, pi);
    return o;
}
void orc_mul_gate(orc_composer *c, uint64_t a, uint64_t b, uint64_t o, const fr_t *q_m, const fr_t *q_o, const fr_t *q_c, const fr_t *pi) {
    fr_t zero = fr_zero();
    push_row(c, a, b, o, c->zero_var, q_m, &zero, &zero, q_o, &zero, q_c, pi);
}
void orc_boolean_gate(orc_composer *c, uint64_t a) {
    fr_t zero = fr_zero(), one = fr_one(), m1 = fr_neg(&one);
    push_row(c, a, a, a, c->zero_var, &one, &zero, &zero, &m1, &zero, &zero, NULL);
}

/* range_gate: restated statement by statement (add_wire closure, padding rule, selector extension, last-gate fix-ups) */
static void range_add_wire(orc_composer *c, uint64_t base, uint64_t i, uint64_t variable) {
    uint64_t gate_index = base + (i / 4);            /* four quads fit into one gate */
    switch (i % 4) {
        case 0: u64vec_push(&c->w[3], variable); perm_add(c, variable, gate_index, 3); break;   /* WireData::Fourth */
        case 1: u64vec_push(&c->w[2], variable); perm_add(c, variable, gate_index, 2); break;   /* WireData::Output */
        case 2: u64vec_push(&c->w[1], variable); perm_add(c, variable, gate_index, 1); break;   /* WireData::Right */
        default: u64vec_push(&c->w[0], variable); perm_add(c, variable, gate_index, 0); break;  /* WireData::Left */
    }
}
void orc_range_gate(orc_composer *c, uint64_t witness, uint64_t num_bits) {
    if (num_bits % 2 != 0 || num_bits < 2 || num_bits > 256) die("range_gate: num_bits must be even, 2..256");
    fr_t value = orc_value_of(c, witness);
    uint8_t bytes[32]; fr_to_bytes(&value, bytes);   /* bits[j] = bit j of the canonical integer (BitIterator8, reversed) */
    uint64_t num_gates = num_bits >> 3;
    if (num_bits % 8 != 0) num_gates += 1;
    uint64_t num_quads = num_gates * 4;
    uint64_t pad = 1 + (((num_quads << 1) - num_bits) >> 1);
    uint64_t used_gates = num_gates + 1;
    uint64_t base = c->n, last_acc = 0;
    fr_t accumulator = fr_zero(), four = fr_from_u64(4), zero = fr_zero(), one = fr_one();
    for (uint64_t i = 0; i < pad; i++) range_add_wire(c, base, i, c->zero_var);
    for (uint64_t i = pad; i <= num_quads; i++) {
        uint64_t bit_index = (num_quads - i) << 1;
        uint64_t q_0 = (bytes[bit_index >> 3] >> (bit_index & 7)) & 1, q_1 = (bytes[(bit_index + 1) >> 3] >> ((bit_index + 1) & 7)) & 1;
        fr_t quad = fr_from_u64(q_0 + 2 * q_1);
        accumulator = fr_mul(&four, &accumulator);
        accumulator = fr_add(&accumulator, &quad);
        last_acc = orc_add_input(c, &accumulator);
        range_add_wire(c, base, i, last_acc);
    }
    for (uint64_t g = 0; g < used_gates; g++)
        for (int s = 0; s < ORC_NSEL; s++) frvec_push(&c->sel[s], s == ORC_QRANGE ? &one : &zero);
    c->n += used_gates;
    c->sel[ORC_QRANGE].p[c->sel[ORC_QRANGE].len - 1] = zero;        /* switch the range selector off on the last gate */
    u64vec_push(&c->w[0], c->zero_var); u64vec_push(&c->w[1], c->zero_var); u64vec_push(&c->w[2], c->zero_var);   /* no permutation entries */
    orc_assert_equal(c, last_acc, witness);
}

orc_composer *orc_composer_new(void) {
    orc_composer *c = (orc_composer *)calloc(1, sizeof(orc_composer));
    if (!c) die("oom");
    map_grow(c);
    c->zero_var = 0;
    fr_t zero = fr_zero();
    c->zero_var = orc_add_witness_to_circuit_description(c, &zero);
    /* add_dummy_constraints(): variables 6, 1, 7, -20 and two rows exercising every arithmetic selector */
    fr_t s6 = fr_from_u64(6), s1 = fr_from_u64(1), s7 = fr_from_u64(7), s20 = fr_from_u64(20), m20 = fr_neg(&s20);
    uint64_t v6 = orc_add_input(c, &s6), v1 = orc_add_input(c, &s1), v7 = orc_add_input(c, &s7), vm20 = orc_add_input(c, &m20);
    fr_t f2 = fr_from_u64(2), f3 = fr_from_u64(3), f4 = fr_from_u64(4), f127 = fr_from_u64(127);
    push_row(c, v6, v7, vm20, v1, &s1, &f2, &f3, &f4, &s1, &f4, NULL);
    push_row(c, vm20, v6, v7, c->zero_var, &s1, &s1, &s1, &s1, &zero, &f127, NULL);
    return c;
}
void orc_composer_free(orc_composer *c) {
    if (!c) return;
    for (int i = 0; i < ORC_NSEL; i++) free(c->sel[i].p);
    for (int i = 0; i < 4; i++) free(c->w[i].p);
    free(c->pi_pos.p); free(c->pi_val.p); free(c->map_key); free(c->map_val);
    free(c->perm_head.p); free(c->perm_next.p); free(c->perm_data.p);
    free(c);
}

uint64_t orc_check(const orc_composer *c, uint64_t *first_bad) {
    uint64_t bad = 0, first = ~(uint64_t)0; size_t pi_i = 0;
    for (size_t i = 0; i < c->n; i++) {
        fr_t a = orc_value_of(c, c->w[0].p[i]), b = orc_value_of(c, c->w[1].p[i]);
        fr_t o = orc_value_of(c, c->w[2].p[i]), d = orc_value_of(c, c->w[3].p[i]);
        fr_t t = fr_mul(&a, &b); t = fr_mul(&c->sel[ORC_QM].p[i], &t);
        fr_t u = fr_mul(&c->sel[ORC_QL].p[i], &a); t = fr_add(&t, &u);
        u = fr_mul(&c->sel[ORC_QR].p[i], &b); t = fr_add(&t, &u);
        u = fr_mul(&c->sel[ORC_QO].p[i], &o); t = fr_add(&t, &u);
        u = fr_mul(&c->sel[ORC_Q4].p[i], &d); t = fr_add(&t, &u);
        if (pi_i < c->pi_pos.len && c->pi_pos.p[pi_i] == i) { t = fr_add(&t, &c->pi_val.p[pi_i]); pi_i++; }
        t = fr_add(&t, &c->sel[ORC_QC].p[i]);
        t = fr_mul(&c->sel[ORC_QARITH].p[i], &t);
        if (!fr_is_zero(&c->sel[ORC_QRANGE].p[i])) {   /* q_range * (delta(c-4d) + delta(b-4c) + delta(a-4b) + delta(d_next-4a)) */
            fr_t dn = orc_value_of(c, c->w[3].p[(i + 1) % c->n]);
            const fr_t *hi[4] = {&o, &b, &a, &dn}, *lo[4] = {&d, &o, &b, &a};
            fr_t four = fr_from_u64(4), one = fr_one(), sum = fr_zero();
            for (int k = 0; k < 4; k++) {
                fr_t f = fr_mul(&four, lo[k]); f = fr_sub(hi[k], &f);
                fr_t g = f, p = f;
                for (int j = 0; j < 3; j++) { g = fr_sub(&g, &one); p = fr_mul(&p, &g); }   /* f (f-1) (f-2) (f-3) */
                sum = fr_add(&sum, &p);
            }
            sum = fr_mul(&c->sel[ORC_QRANGE].p[i], &sum);
            t = fr_add(&t, &sum);
        }
        if (!fr_is_zero(&t)) { if (!bad) first = i; bad++; }
    }
    if (first_bad) *first_bad = first;
    return bad;
}
void orc_dense_pi(const orc_composer *c, fr_t *out) {
    for (size_t i = 0; i < c->n; i++) out[i] = fr_zero();
    for (size_t k = 0; k < c->pi_pos.len; k++) out[c->pi_pos.p[k]] = c->pi_val.p[k];
}
