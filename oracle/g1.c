/*
 * oracle/g1.c -- CPU restatement of the commitment step of Prover::prove (second half):
 *     w_l_poly_commit = commit_key.commit(&w_l_poly)  =  msm_variable_base(&powers_of_g, &poly.coeffs)
 * over the BLS12-381 group G1.
 *
 * TEST INFRASTRUCTURE ONLY (see fr.h).  PARITY STATUS: "parity unpinned".  The algorithms live in third-party crates that
 * /root/reference only names in Cargo.toml:20 (dusk-plonk 0.8 -> dusk-bls12_381: fp.rs, g1.rs, multiscalar_mul.rs;
 * dusk-plonk src/commitment_scheme/kzg10/{key.rs,srs.rs}); the reference reaches them through prover.prove
 * (/root/reference/tests/range_gadgets_tests.rs:90-91).  Restated here:
 *   * Fp: 6 x u64 little-endian limbs of a*2^384 mod p, fully reduced, schoolbook multiply + Montgomery reduction;
 *   * G1: y^2 = x^3 + 4; Jacobian coordinates internally (the crates use complete projective formulas; the affine result of a
 *     group operation does not depend on the coordinate system);
 *   * msm_variable_base: the crate's serial Pippenger -- window c = 3 for fewer than 32 scalars, else ln(n) + 2; 2^c - 1
 *     buckets per window; scalars equal to one are added directly in window 0; running-sum bucket reduction; windows
 *     combined from the top with c doublings each;
 *   * PublicParameters::setup: powers_of_g[i] = beta^i * g.
 * Pins: the field/curve constants are recomputed from the BLS parameter x = -0xd201000000010000 and checked (generator on
 * the curve, of order q) in tests/test_oracle_g1.py; every operation is cross-checked against the independent affine
 * big-int model in oracle/pymodel.py.
 */
#include "fr.h"
#include <stdlib.h>

typedef struct { uint64_t l[6]; } fp_t;
typedef struct { fp_t x, y; int inf; } g1_affine_t;          /* API form: canonical Montgomery limbs + infinity flag */
typedef struct { fp_t x, y, z; } g1_jac_t;                   /* z == 0: infinity */

static const fp_t FP_P  = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const fp_t FP_R  = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};
static const fp_t FP_R2 = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL, 0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
#define FP_INV 0x89f3fffcfffcfffdULL
static const fp_t G1_GEN_X = {{0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL, 0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL}};
static const fp_t G1_GEN_Y = {{0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL, 0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL}};

static int fp_is_zero(const fp_t *a) { uint64_t o = 0; for (int i = 0; i < 6; i++) o |= a->l[i]; return o == 0; }
static int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a, b, sizeof(fp_t)) == 0; }
static fp_t fp_zero(void) { fp_t z; memset(&z, 0, sizeof z); return z; }

/* r in [0, 2p) -> r mod p */
static fp_t fp_sub_cond(const fp_t *r) {
    fp_t d; uint64_t bw = 0;
    for (int i = 0; i < 6; i++) d.l[i] = fr_sbb(r->l[i], FP_P.l[i], &bw);
    return bw ? *r : d;
}
static fp_t fp_sub(const fp_t *a, const fp_t *b) {
    fp_t d; uint64_t bw = 0, c = 0;
    for (int i = 0; i < 6; i++) d.l[i] = fr_sbb(a->l[i], b->l[i], &bw);
    for (int i = 0; i < 6; i++) d.l[i] = fr_adc(d.l[i], FP_P.l[i] & bw, &c);
    return d;
}
static fp_t fp_add(const fp_t *a, const fp_t *b) {
    fp_t s; uint64_t c = 0;
    for (int i = 0; i < 6; i++) s.l[i] = fr_adc(a->l[i], b->l[i], &c);
    return fp_sub_cond(&s);                        /* p < 2^381: the sum fits 6 limbs */
}
static fp_t fp_neg(const fp_t *a) { fp_t z = fp_zero(); return fp_is_zero(a) ? z : fp_sub(&z, a); }
static fp_t fp_mul(const fp_t *a, const fp_t *b) {
    uint64_t t[13]; memset(t, 0, sizeof t);
    for (int i = 0; i < 6; i++) {                  /* schoolbook product */
        uint64_t carry = 0;
        for (int j = 0; j < 6; j++) t[i + j] = fr_mac(t[i + j], a->l[i], b->l[j], &carry);
        t[i + 6] = carry;
    }
    for (int i = 0; i < 6; i++) {                  /* montgomery_reduce */
        const uint64_t k = t[i] * FP_INV; uint64_t carry = 0;
        (void)fr_mac(t[i], k, FP_P.l[0], &carry);
        for (int j = 1; j < 6; j++) t[i + j] = fr_mac(t[i + j], k, FP_P.l[j], &carry);
        uint64_t c2 = 0;
        t[i + 6] = fr_adc(t[i + 6], carry, &c2);
        for (int j = i + 7; j < 13; j++) t[j] = fr_adc(t[j], 0, &c2);
    }
    fp_t r; for (int i = 0; i < 6; i++) r.l[i] = t[i + 6];   /* a, b < p => value < 2p < 2^382: t[12] == 0 */
    return fp_sub_cond(&r);
}
static fp_t fp_sqr(const fp_t *a) { return fp_mul(a, a); }
static fp_t fp_from_raw(const fp_t *raw) { return fp_mul(raw, &FP_R2); }
static fp_t fp_to_raw(const fp_t *a) { fp_t one = fp_zero(); one.l[0] = 1; return fp_mul(a, &one); }
static fp_t fp_inv(const fp_t *a) {                /* a^(p-2) */
    fp_t e = FP_P; e.l[0] -= 2;
    fp_t res = FP_R;
    for (int i = 5; i >= 0; i--) for (int b = 63; b >= 0; b--) {
        res = fp_sqr(&res);
        if ((e.l[i] >> b) & 1) res = fp_mul(&res, a);
    }
    return res;
}

/* ---- G1 ---- */
static g1_jac_t jac_inf(void) { g1_jac_t r; memset(&r, 0, sizeof r); r.x = FP_R; r.y = FP_R; return r; }
static g1_jac_t jac_from_affine(const g1_affine_t *a) { g1_jac_t r; if (a->inf) return jac_inf(); r.x = a->x; r.y = a->y; r.z = FP_R; return r; }
static g1_jac_t jac_double(const g1_jac_t *p) {
    if (fp_is_zero(&p->z)) return *p;
    fp_t a = fp_sqr(&p->x), b = fp_sqr(&p->y), c = fp_sqr(&b);
    fp_t t = fp_add(&p->x, &b); t = fp_sqr(&t); t = fp_sub(&t, &a); t = fp_sub(&t, &c);
    fp_t d = fp_add(&t, &t);
    fp_t e = fp_add(&a, &a); e = fp_add(&e, &a);
    fp_t f = fp_sqr(&e);
    g1_jac_t r;
    fp_t d2 = fp_add(&d, &d);
    r.x = fp_sub(&f, &d2);
    fp_t c8 = fp_add(&c, &c); c8 = fp_add(&c8, &c8); c8 = fp_add(&c8, &c8);
    fp_t dx = fp_sub(&d, &r.x);
    r.y = fp_mul(&e, &dx); r.y = fp_sub(&r.y, &c8);
    r.z = fp_mul(&p->y, &p->z); r.z = fp_add(&r.z, &r.z);
    return r;
}
static g1_jac_t jac_add(const g1_jac_t *p, const g1_jac_t *q) {
    if (fp_is_zero(&p->z)) return *q;
    if (fp_is_zero(&q->z)) return *p;
    fp_t z1z1 = fp_sqr(&p->z), z2z2 = fp_sqr(&q->z);
    fp_t u1 = fp_mul(&p->x, &z2z2), u2 = fp_mul(&q->x, &z1z1);
    fp_t s1 = fp_mul(&p->y, &q->z); s1 = fp_mul(&s1, &z2z2);
    fp_t s2 = fp_mul(&q->y, &p->z); s2 = fp_mul(&s2, &z1z1);
    if (fp_eq(&u1, &u2)) { if (fp_eq(&s1, &s2)) return jac_double(p); return jac_inf(); }
    fp_t h = fp_sub(&u2, &u1), r = fp_sub(&s2, &s1);
    fp_t hh = fp_sqr(&h), hhh = fp_mul(&hh, &h), v = fp_mul(&u1, &hh);
    g1_jac_t o;
    fp_t v2 = fp_add(&v, &v);
    o.x = fp_sqr(&r); o.x = fp_sub(&o.x, &hhh); o.x = fp_sub(&o.x, &v2);
    fp_t vx = fp_sub(&v, &o.x), s1h = fp_mul(&s1, &hhh);
    o.y = fp_mul(&r, &vx); o.y = fp_sub(&o.y, &s1h);
    o.z = fp_mul(&p->z, &q->z); o.z = fp_mul(&o.z, &h);
    return o;
}
static g1_jac_t jac_add_mixed(const g1_jac_t *p, const g1_affine_t *q) { g1_jac_t j = jac_from_affine(q); return jac_add(p, &j); }
static g1_affine_t jac_to_affine(const g1_jac_t *p) {
    g1_affine_t a; memset(&a, 0, sizeof a);
    if (fp_is_zero(&p->z)) { a.inf = 1; return a; }
    fp_t zi = fp_inv(&p->z), zi2 = fp_sqr(&zi), zi3 = fp_mul(&zi2, &zi);
    a.x = fp_mul(&p->x, &zi2); a.y = fp_mul(&p->y, &zi3); a.inf = 0;
    return a;
}
static g1_jac_t jac_mul(const g1_jac_t *p, const fr_t *k_mont) {       /* k in Montgomery form, like every BlsScalar */
    const fr_t k = fr_reduce(k_mont);
    g1_jac_t acc = jac_inf();
    for (int i = 3; i >= 0; i--) for (int b = 63; b >= 0; b--) {
        acc = jac_double(&acc);
        if ((k.l[i] >> b) & 1) acc = jac_add(&acc, p);
    }
    return acc;
}

/* ln_without_floats of the crate: floor(log2(n)) * 69 / 100 */
static unsigned ln_without_floats(uint64_t a) { unsigned l = 0; while ((a >> (l + 1)) != 0) l++; return l * 69 / 100; }

/* msm_variable_base(points, scalars) */
void orc_g1_msm(uint64_t n, const g1_affine_t *points, const fr_t *scalars, g1_affine_t *out) {
    const unsigned c = n < 32 ? 3 : ln_without_floats(n) + 2;
    const unsigned num_bits = 255;
    const fr_t one = fr_one();
    const unsigned n_windows = (num_bits + c - 1) / c;
    g1_jac_t *sums = malloc(n_windows * sizeof(g1_jac_t));
    g1_jac_t *buckets = malloc(((size_t)1 << c) * sizeof(g1_jac_t));
    fr_t *canon = malloc((n ? n : 1) * sizeof(fr_t));
    for (uint64_t i = 0; i < n; i++) canon[i] = fr_reduce(&scalars[i]);
    for (unsigned w = 0; w < n_windows; w++) {
        const unsigned w_start = w * c;
        g1_jac_t res = jac_inf();
        for (size_t b = 0; b + 1 < ((size_t)1 << c); b++) buckets[b] = jac_inf();
        for (uint64_t i = 0; i < n; i++) {
            if (fr_is_zero(&scalars[i])) continue;
            if (fr_eq(&scalars[i], &one)) { if (w_start == 0) res = jac_add_mixed(&res, &points[i]); continue; }
            fr_t s = canon[i];
            fr_divn(&s, w_start);
            const uint64_t d = s.l[0] % ((uint64_t)1 << c);
            if (d) buckets[d - 1] = jac_add_mixed(&buckets[d - 1], &points[i]);
        }
        g1_jac_t running = jac_inf();
        for (size_t b = ((size_t)1 << c) - 1; b-- > 0;) { running = jac_add(&running, &buckets[b]); res = jac_add(&res, &running); }
        sums[w] = res;
    }
    g1_jac_t total = jac_inf();
    for (unsigned w = n_windows; w-- > 1;) {
        total = jac_add(&total, &sums[w]);
        for (unsigned k = 0; k < c; k++) total = jac_double(&total);
    }
    total = jac_add(&total, &sums[0]);
    *out = jac_to_affine(&total);
    free(sums); free(buckets); free(canon);
}

/* ---- flat API (ctypes) ---- */
void orc_fp_from_raw(const fp_t *raw, fp_t *out) { *out = fp_from_raw(raw); }
void orc_fp_to_raw(const fp_t *a, fp_t *out) { *out = fp_to_raw(a); }
void orc_fp_mul(const fp_t *a, const fp_t *b, fp_t *out) { *out = fp_mul(a, b); }
void orc_fp_add(const fp_t *a, const fp_t *b, fp_t *out) { *out = fp_add(a, b); }
void orc_fp_sub(const fp_t *a, const fp_t *b, fp_t *out) { *out = fp_sub(a, b); }
void orc_fp_neg(const fp_t *a, fp_t *out) { *out = fp_neg(a); }
void orc_fp_inv(const fp_t *a, fp_t *out) { *out = fp_inv(a); }
void orc_g1_generator(g1_affine_t *out) { out->x = G1_GEN_X; out->y = G1_GEN_Y; out->inf = 0; }
int orc_g1_on_curve(const g1_affine_t *a) {
    if (a->inf) return 1;
    fp_t y2 = fp_sqr(&a->y), x3 = fp_sqr(&a->x); x3 = fp_mul(&x3, &a->x);
    fp_t four = fp_add(&FP_R, &FP_R); four = fp_add(&four, &four);
    x3 = fp_add(&x3, &four);
    return fp_eq(&y2, &x3);
}
void orc_g1_add(const g1_affine_t *a, const g1_affine_t *b, g1_affine_t *out) {
    g1_jac_t ja = jac_from_affine(a); g1_jac_t s = jac_add_mixed(&ja, b); *out = jac_to_affine(&s);
}
void orc_g1_mul(const g1_affine_t *a, const fr_t *k, g1_affine_t *out) {
    g1_jac_t ja = jac_from_affine(a); g1_jac_t s = jac_mul(&ja, k); *out = jac_to_affine(&s);
}
/* powers_of_g[i] = beta^i * g, i < n */
void orc_srs_powers(const fr_t *beta, const g1_affine_t *g, uint64_t n, g1_affine_t *out) {
    fr_t e = fr_one();
    g1_jac_t jg = jac_from_affine(g);
    for (uint64_t i = 0; i < n; i++) { g1_jac_t s = jac_mul(&jg, &e); out[i] = jac_to_affine(&s); e = fr_mul(&e, beta); }
}
