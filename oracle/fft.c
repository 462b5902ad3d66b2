/*
 * oracle/fft.c -- CPU restatement of the step that follows the gadget hot path inside `Prover::prove`
 * (first half): wire columns -> zero-padded scalar vectors -> inverse FFT over the
 * evaluation domain -> coefficient vectors of w_l(X), w_r(X), w_o(X), w_4(X).
 *
 * TEST INFRASTRUCTURE ONLY (see fr.h).  PARITY STATUS: "parity unpinned".  The algorithm lives in the third-party
 * crate dusk-plonk 0.8 (`src/fft/domain.rs`: EvaluationDomain::new / fft / ifft / serial_fft; `src/proof_system/
 * prover.rs`: "Convert Variables to BlsScalars padding them to the correct domain size", `domain.ifft(w_l_scalar)`),
 * which /root/reference only names in Cargo.toml:20; it is exercised by the reference through
 * /root/reference/tests/range_gadgets_tests.rs:90-91 and tests/scalar_gadgets_tests.rs (prover.prove).  This file
 * restates the published algorithm:
 *   * domain size = circuit_size.next_power_of_two(), log_size = trailing_zeros, at most TWO_ADICITY = 32;
 *   * group_gen = ROOT_OF_UNITY squared (32 - log_size) times, ROOT_OF_UNITY = 7^((q-1)/2^32);
 *   * serial_fft: bit-reversal permutation, then log_size rounds of radix-2 decimation-in-time butterflies with
 *     w_m = omega^(n/2m), w stepping by w_m inside a half block;
 *   * ifft = the same with group_gen^-1, followed by a multiplication of every element by size^-1.
 * Pins: the ROOT_OF_UNITY limbs below are the crate's constant as recalled AND are recomputed from 7^((q-1)/2^32)
 * by tests/test_oracle_fft.py; the transform itself is checked against the O(n^2) big-int DFT of oracle/pymodel.py
 * (the DFT of a vector is unique, so any correct algorithm yields the same fully reduced Montgomery limbs).
 */
#include "composer.h"
#include <stdlib.h>

static const fr_t FR_ROOT_OF_UNITY = {{0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL, 0x5bf3adda19e9b27bULL}};
#define FR_TWO_ADICITY 32

/* EvaluationDomain::new: generator of the size-2^log_n subgroup */
static fr_t domain_group_gen(unsigned log_n) {
    fr_t g = FR_ROOT_OF_UNITY;
    for (unsigned i = log_n; i < FR_TWO_ADICITY; i++) g = fr_square(&g);
    return g;
}

static uint32_t bitreverse(uint32_t n, unsigned l) {
    uint32_t r = 0;
    for (unsigned i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}

/* serial_fft(a, omega, log_n) */
static void serial_fft(fr_t *a, uint64_t n, const fr_t *omega, unsigned log_n) {
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t rk = bitreverse((uint32_t)k, log_n);
        if (k < rk) { fr_t t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    uint64_t m = 1;
    for (unsigned s = 0; s < log_n; s++) {
        const uint64_t e[4] = {n / (2 * m), 0, 0, 0};
        const fr_t w_m = fr_pow(omega, e);
        for (uint64_t k = 0; k < n; k += 2 * m) {
            fr_t w = fr_one();
            for (uint64_t j = 0; j < m; j++) {
                fr_t t = fr_mul(&a[k + j + m], &w);
                fr_t tmp = fr_sub(&a[k + j], &t);
                a[k + j + m] = tmp;
                a[k + j] = fr_add(&a[k + j], &t);
                w = fr_mul(&w, &w_m);
            }
        }
        m *= 2;
    }
}

/* EvaluationDomain::fft_in_place / ifft_in_place on a vector already resized to 2^log_n.  Returns 0, or -1 for log_n > 32. */
int orc_fft(fr_t *a, unsigned log_n, int inverse) {
    if (log_n > FR_TWO_ADICITY) return -1;
    const uint64_t n = 1ull << log_n;
    fr_t gen = domain_group_gen(log_n);
    if (!inverse) { serial_fft(a, n, &gen, log_n); return 0; }
    fr_t gen_inv, size_inv, size = fr_from_u64(n);
    fr_invert(&gen, &gen_inv);
    fr_invert(&size, &size_inv);
    serial_fft(a, n, &gen_inv, log_n);
    for (uint64_t i = 0; i < n; i++) a[i] = fr_mul(&a[i], &size_inv);
    return 0;
}

void orc_domain_group_gen(unsigned log_n, fr_t *out) { *out = domain_group_gen(log_n); }

/* log2 of EvaluationDomain::new(circuit_size).size() */
unsigned orc_domain_log_size(uint64_t circuit_size) {
    unsigned l = 0;
    while ((1ull << l) < circuit_size) l++;
    return l;
}

/* Prover round 1, before blinding/commitment: out[w * size + i] = coefficient i of the wire polynomial w
 * (w = 0..3: w_l, w_r, w_o, w_4), size = 2^orc_domain_log_size(n).  Returns log2(size). */
unsigned orc_wire_polynomials(const orc_composer *c, fr_t *out) {
    const unsigned log_n = orc_domain_log_size(c->n);
    const uint64_t size = 1ull << log_n;
    for (int w = 0; w < 4; w++) {
        fr_t *col = out + (uint64_t)w * size;
        for (uint64_t i = 0; i < size; i++) col[i] = i < c->n ? orc_value_of(c, c->w[w].p[i]) : fr_zero();   /* to_scalars + pad */
        orc_fft(col, log_n, 1);
    }
    return log_n;
}
