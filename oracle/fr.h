/*
 * oracle/fr.h -- CPU restatement of dusk-bls12_381's `Scalar` (= dusk_plonk::bls12_381::BlsScalar).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by, or executed from the
 * product library (plonk_gadgets_b200/): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at limb/row level.  The arithmetic lives in a third-party crate
 * (dusk-bls12_381, pulled in transitively by `dusk-plonk = "0.8"`, /root/reference/Cargo.toml:20; no
 * Cargo.lock, sources not vendored, no Rust toolchain here), so this file restates the crate's published
 * algorithm: 4 x u64 little-endian limbs of a*2^256 mod q, always fully reduced,
 * schoolbook 4x4 multiply followed by `montgomery_reduce`, Fermat inversion, fixed 256-step `pow`.
 * It is pinned by (i) the constants recomputed from q, (ii) the independent big-int model
 * oracle/pymodel.py, (iii) the reference's verdict-level KATs (tests/test_oracle_kats.py).
 *
 * Reference call sites served: invert (/root/reference/src/scalar.rs:73,:122), pow (range.rs:146),
 * pow_of_2 (range.rs:187), reduce/divn/Ord (range.rs:174-177), to_bytes (range.rs:163),
 * From<u64> (range.rs:130,:146,:152), + - * neg (range.rs:62,:69,:87,:94,:102,:152; scalar.rs:47,:113,:121,:126).
 */
#ifndef ORACLE_FR_H
#define ORACLE_FR_H

#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct { uint64_t l[4]; } fr_t;   /* Montgomery form, LE limbs == BlsScalar([u64;4]) */

static const fr_t FR_MODULUS = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
static const fr_t FR_R  = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};
static const fr_t FR_R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};
static const fr_t FR_R3 = {{0xc62c1807439b73afULL, 0x1b3e0d188cf06990ULL, 0x73d13c71c7b5f418ULL, 0x6e2a5bb9c8db33e9ULL}};
#define FR_INV 0xfffffffeffffffffULL   /* -q^{-1} mod 2^64 */

static inline uint64_t fr_adc(uint64_t a, uint64_t b, uint64_t *carry) {
    u128 t = (u128)a + b + *carry; *carry = (uint64_t)(t >> 64); return (uint64_t)t;
}
static inline uint64_t fr_sbb(uint64_t a, uint64_t b, uint64_t *borrow) {
    u128 t = (u128)a - b - (*borrow >> 63); *borrow = (uint64_t)(t >> 64); return (uint64_t)t;
}
static inline uint64_t fr_mac(uint64_t a, uint64_t b, uint64_t c, uint64_t *carry) {
    u128 t = (u128)a + (u128)b * c + *carry; *carry = (uint64_t)(t >> 64); return (uint64_t)t;
}

static inline fr_t fr_zero(void) { fr_t z = {{0, 0, 0, 0}}; return z; }
static inline fr_t fr_one(void) { return FR_R; }
static inline int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, sizeof(fr_t)) == 0; }
static inline int fr_is_zero(const fr_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }

/* a - b, adding q back when the subtraction borrowed */
static inline fr_t fr_sub(const fr_t *a, const fr_t *b) {
    uint64_t bw = 0, c = 0; fr_t d;
    d.l[0] = fr_sbb(a->l[0], b->l[0], &bw); d.l[1] = fr_sbb(a->l[1], b->l[1], &bw);
    d.l[2] = fr_sbb(a->l[2], b->l[2], &bw); d.l[3] = fr_sbb(a->l[3], b->l[3], &bw);
    d.l[0] = fr_adc(d.l[0], FR_MODULUS.l[0] & bw, &c); d.l[1] = fr_adc(d.l[1], FR_MODULUS.l[1] & bw, &c);
    d.l[2] = fr_adc(d.l[2], FR_MODULUS.l[2] & bw, &c); d.l[3] = fr_adc(d.l[3], FR_MODULUS.l[3] & bw, &c);
    return d;
}
static inline fr_t fr_add(const fr_t *a, const fr_t *b) {
    uint64_t c = 0; fr_t d;
    d.l[0] = fr_adc(a->l[0], b->l[0], &c); d.l[1] = fr_adc(a->l[1], b->l[1], &c);
    d.l[2] = fr_adc(a->l[2], b->l[2], &c); d.l[3] = fr_adc(a->l[3], b->l[3], &c);
    return fr_sub(&d, &FR_MODULUS);   /* q < 2^255 so no carry is lost */
}
static inline fr_t fr_neg(const fr_t *a) {
    uint64_t bw = 0; fr_t d;
    d.l[0] = fr_sbb(FR_MODULUS.l[0], a->l[0], &bw); d.l[1] = fr_sbb(FR_MODULUS.l[1], a->l[1], &bw);
    d.l[2] = fr_sbb(FR_MODULUS.l[2], a->l[2], &bw); d.l[3] = fr_sbb(FR_MODULUS.l[3], a->l[3], &bw);
    uint64_t mask = fr_is_zero(a) ? 0 : ~(uint64_t)0;
    d.l[0] &= mask; d.l[1] &= mask; d.l[2] &= mask; d.l[3] &= mask;
    return d;
}

/* montgomery_reduce(r0..r7): four rounds of k = r_i * INV; r += k*q << 64i; then conditional subtract */
static inline fr_t fr_montgomery_reduce(uint64_t r0, uint64_t r1, uint64_t r2, uint64_t r3,
                                        uint64_t r4, uint64_t r5, uint64_t r6, uint64_t r7) {
    const uint64_t *m = FR_MODULUS.l; uint64_t k, carry, carry2 = 0;
    k = r0 * FR_INV; carry = 0;
    (void)fr_mac(r0, k, m[0], &carry); r1 = fr_mac(r1, k, m[1], &carry); r2 = fr_mac(r2, k, m[2], &carry);
    r3 = fr_mac(r3, k, m[3], &carry); r4 = fr_adc(r4, 0, &carry); carry2 = carry;
    k = r1 * FR_INV; carry = 0;
    (void)fr_mac(r1, k, m[0], &carry); r2 = fr_mac(r2, k, m[1], &carry); r3 = fr_mac(r3, k, m[2], &carry);
    r4 = fr_mac(r4, k, m[3], &carry); r5 = fr_adc(r5, carry2, &carry); carry2 = carry;
    k = r2 * FR_INV; carry = 0;
    (void)fr_mac(r2, k, m[0], &carry); r3 = fr_mac(r3, k, m[1], &carry); r4 = fr_mac(r4, k, m[2], &carry);
    r5 = fr_mac(r5, k, m[3], &carry); r6 = fr_adc(r6, carry2, &carry); carry2 = carry;
    k = r3 * FR_INV; carry = 0;
    (void)fr_mac(r3, k, m[0], &carry); r4 = fr_mac(r4, k, m[1], &carry); r5 = fr_mac(r5, k, m[2], &carry);
    r6 = fr_mac(r6, k, m[3], &carry); r7 = fr_adc(r7, carry2, &carry);
    fr_t t = {{r4, r5, r6, r7}};
    return fr_sub(&t, &FR_MODULUS);
}

static inline fr_t fr_mul(const fr_t *a, const fr_t *b) {
    uint64_t c, r0, r1, r2, r3, r4, r5, r6, r7;
    c = 0; r0 = fr_mac(0, a->l[0], b->l[0], &c); r1 = fr_mac(0, a->l[0], b->l[1], &c);
    r2 = fr_mac(0, a->l[0], b->l[2], &c); r3 = fr_mac(0, a->l[0], b->l[3], &c); r4 = c;
    c = 0; r1 = fr_mac(r1, a->l[1], b->l[0], &c); r2 = fr_mac(r2, a->l[1], b->l[1], &c);
    r3 = fr_mac(r3, a->l[1], b->l[2], &c); r4 = fr_mac(r4, a->l[1], b->l[3], &c); r5 = c;
    c = 0; r2 = fr_mac(r2, a->l[2], b->l[0], &c); r3 = fr_mac(r3, a->l[2], b->l[1], &c);
    r4 = fr_mac(r4, a->l[2], b->l[2], &c); r5 = fr_mac(r5, a->l[2], b->l[3], &c); r6 = c;
    c = 0; r3 = fr_mac(r3, a->l[3], b->l[0], &c); r4 = fr_mac(r4, a->l[3], b->l[1], &c);
    r5 = fr_mac(r5, a->l[3], b->l[2], &c); r6 = fr_mac(r6, a->l[3], b->l[3], &c); r7 = c;
    return fr_montgomery_reduce(r0, r1, r2, r3, r4, r5, r6, r7);
}
static inline fr_t fr_square(const fr_t *a) { return fr_mul(a, a); }

/* From<u64>: Scalar([v,0,0,0]) * R2 */
static inline fr_t fr_from_u64(uint64_t v) { fr_t t = {{v, 0, 0, 0}}; return fr_mul(&t, &FR_R2); }
/* from_raw: raw * R2 */
static inline fr_t fr_from_raw(const uint64_t v[4]) { fr_t t = {{v[0], v[1], v[2], v[3]}}; return fr_mul(&t, &FR_R2); }

/* reduce(): montgomery_reduce(l0..l3,0,0,0,0) -- a Scalar whose RAW limbs are the canonical integer */
static inline fr_t fr_reduce(const fr_t *a) { return fr_montgomery_reduce(a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0); }

/* to_bytes(): canonical little-endian 32 bytes */
static inline void fr_to_bytes(const fr_t *a, uint8_t out[32]) {
    fr_t c = fr_reduce(a);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c.l[i] >> (8 * j));
}
/* from_bytes(): returns 0 and leaves *out untouched when the encoding is >= q */
static inline int fr_from_bytes(const uint8_t in[32], fr_t *out) {
    uint64_t v[4];
    for (int i = 0; i < 4; i++) { v[i] = 0; for (int j = 0; j < 8; j++) v[i] |= (uint64_t)in[8 * i + j] << (8 * j); }
    uint64_t bw = 0;
    (void)fr_sbb(v[0], FR_MODULUS.l[0], &bw); (void)fr_sbb(v[1], FR_MODULUS.l[1], &bw);
    (void)fr_sbb(v[2], FR_MODULUS.l[2], &bw); (void)fr_sbb(v[3], FR_MODULUS.l[3], &bw);
    if (!(bw >> 63)) return 0;
    *out = fr_from_raw(v); return 1;
}
/* from_bytes_wide(): 64 LE bytes -> d0*R2 + d1*R3 */
static inline fr_t fr_from_bytes_wide(const uint8_t in[64]) {
    uint64_t v[8];
    for (int i = 0; i < 8; i++) { v[i] = 0; for (int j = 0; j < 8; j++) v[i] |= (uint64_t)in[8 * i + j] << (8 * j); }
    fr_t d0 = {{v[0], v[1], v[2], v[3]}}, d1 = {{v[4], v[5], v[6], v[7]}};
    fr_t a = fr_mul(&d0, &FR_R2), b = fr_mul(&d1, &FR_R3);
    return fr_add(&a, &b);
}

/* divn(n): shift the RAW limbs right by n bits (n < 256) */
static inline void fr_divn(fr_t *a, uint32_t n) {
    if (n >= 256) { *a = fr_zero(); return; }
    while (n >= 64) { uint64_t t = 0; for (int i = 3; i >= 0; i--) { uint64_t x = a->l[i]; a->l[i] = t; t = x; } n -= 64; }
    if (n > 0) { uint64_t t = 0; for (int i = 3; i >= 0; i--) { uint64_t t2 = a->l[i] << (64 - n); a->l[i] = (a->l[i] >> n) | t; t = t2; } }
}
/* Ord: compares RAW limbs, most significant first */
static inline int fr_cmp_raw(const fr_t *a, const fr_t *b) {
    for (int i = 3; i >= 0; i--) { if (a->l[i] < b->l[i]) return -1; if (a->l[i] > b->l[i]) return 1; }
    return 0;
}

/* pow(&[u64;4]): fixed 256 iterations MSB->LSB; res = res^2; tmp = res*self; select.  512 multiplications. */
static inline fr_t fr_pow(const fr_t *self, const uint64_t by[4]) {
    fr_t res = fr_one();
    for (int e = 3; e >= 0; e--) for (int i = 63; i >= 0; i--) {
        res = fr_square(&res);
        fr_t tmp = fr_mul(&res, self);
        uint64_t mask = (uint64_t)0 - ((by[e] >> i) & 1);
        for (int j = 0; j < 4; j++) res.l[j] = (res.l[j] & ~mask) | (tmp.l[j] & mask);
    }
    return res;
}
/* pow_of_2(by): 2^by mod q with a fixed 64-iteration ladder */
static inline fr_t fr_pow_of_2(uint64_t by) {
    fr_t two = fr_from_u64(2), res = fr_one();
    for (int i = 63; i >= 0; i--) {
        res = fr_square(&res);
        fr_t tmp = fr_mul(&res, &two);
        uint64_t mask = (uint64_t)0 - ((by >> i) & 1);
        for (int j = 0; j < 4; j++) res.l[j] = (res.l[j] & ~mask) | (tmp.l[j] & mask);
    }
    return res;
}

/* invert(): Fermat x^(q-2), 4-bit fixed window (the crate uses a hand-made addition chain of similar
 * length; the result -- the unique fully reduced Montgomery form of x^-1 -- is identical).
 * Returns 0 (and *out = 0) iff x == 0 (CtOption::None). */
static inline int fr_invert(const fr_t *x, fr_t *out) {
    if (fr_is_zero(x)) { *out = fr_zero(); return 0; }
    static const uint64_t e[4] = {0xfffffffeffffffffULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}; /* q-2 */
    fr_t tab[16]; tab[0] = fr_one(); tab[1] = *x;
    for (int i = 2; i < 16; i++) tab[i] = fr_mul(&tab[i - 1], x);
    fr_t res = fr_one();
    for (int w = 63; w >= 0; w--) {
        if (w != 63) { res = fr_square(&res); res = fr_square(&res); res = fr_square(&res); res = fr_square(&res); }
        unsigned nib = (unsigned)((e[w / 16] >> (4 * (w % 16))) & 0xf);
        if (nib) res = fr_mul(&res, &tab[nib]);
    }
    *out = res; return 1;
}

#endif /* ORACLE_FR_H */
