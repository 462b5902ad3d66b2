/* oracle/gadgets.c -- see gadgets.h.  TEST INFRASTRUCTURE ONLY. */
#include "gadgets.h"
#include <stdlib.h>

static int g_mode = ORC_FAITHFUL;
static fr_t g_pow2[256];
static int g_pow2_ready = 0;

void orc_set_mode(int mode) {
    g_mode = mode;
    if (mode == ORC_FAST && !g_pow2_ready) {
        fr_t two = fr_from_u64(2); g_pow2[0] = fr_one();
        for (int i = 1; i < 256; i++) g_pow2[i] = fr_mul(&g_pow2[i - 1], &two);
        g_pow2_ready = 1;
    }
}

/* BlsScalar::from(2).pow(&[power, 0, 0, 0])  -- range.rs:146 */
static fr_t two_pow(uint64_t power) {
    if (g_mode == ORC_FAST && g_pow2_ready && power < 256) return g_pow2[power];
    fr_t two = fr_from_u64(2); uint64_t e[4] = {power, 0, 0, 0};
    return fr_pow(&two, e);
}

orc_allocated_scalar orc_allocate(orc_composer *c, const fr_t *scalar) {
    orc_allocated_scalar a; a.var = orc_add_input(c, scalar); a.scalar = *scalar; return a;
}

/* ------------------------------------------------------------------ scalar.rs */
uint64_t orc_conditionally_select_zero(orc_composer *c, uint64_t x, uint64_t select) {
    fr_t one = fr_one(), zero = fr_zero();
    return orc_mul(c, &one, x, select, &zero, NULL);                                   /* scalar.rs:26 */
}

uint64_t orc_conditionally_select_one(orc_composer *c, uint64_t y, uint64_t selector) {
    fr_t one = fr_one(), zero = fr_zero(), m1 = fr_neg(&one);
    uint64_t v_one = orc_add_witness_to_circuit_description(c, &one);                  /* :41 */
    uint64_t selector_y = orc_mul(c, &one, y, selector, &zero, NULL);                  /* :43 */
    uint64_t one_min_selector = orc_add(c, &one, v_one, &m1, selector, &zero, NULL);   /* :45-50 */
    return orc_add(c, &one, selector_y, &one, one_min_selector, &zero, NULL);          /* :53-58 */
}

int orc_is_non_zero(orc_composer *c, uint64_t var, fr_t value_assigned) {
    uint64_t var_assigned = orc_add_input(c, &value_assigned);                         /* :69 */
    orc_assert_equal(c, var, var_assigned);                                            /* :71 */
    fr_t inverse;
    if (!fr_invert(&value_assigned, &inverse)) return ORC_ERR_NON_EXISTING_INVERSE;    /* :73-80 */
    uint64_t inv = orc_add_input(c, &inverse);                                         /* :77 */
    fr_t one = fr_one(), zero = fr_zero(), m1 = fr_neg(&one);
    uint64_t v_one = orc_add_witness_to_circuit_description(c, &one);                  /* :83 */
    orc_poly_gate(c, var, inv, v_one, &one, &zero, &zero, &m1, &zero, NULL);           /* :84-94 */
    return ORC_OK;
}

uint64_t orc_maybe_equal(orc_composer *c, orc_allocated_scalar a, orc_allocated_scalar b) {
    fr_t one = fr_one(), zero = fr_zero(), m1 = fr_neg(&one);
    uint64_t u = orc_add(c, &one, a.var, &m1, b.var, &zero, NULL);                     /* :111-117 */
    fr_t u_scalar = fr_sub(&a.scalar, &b.scalar);                                      /* :121 */
    fr_t u_inv; (void)fr_invert(&u_scalar, &u_inv);                                    /* :122 unwrap_or(zero) */
    uint64_t z = orc_add_input(c, &u_inv);                                             /* :123 */
    uint64_t y = orc_mul(c, &m1, z, u, &one, NULL);                                    /* :126 */
    orc_mul_gate(c, y, u, u, &one, &zero, &zero, NULL);                                /* :129-138 */
    return y;
}

/* ------------------------------------------------------------------ range.rs */
void orc_scalar_to_bits(const fr_t *scalar, uint8_t out[256]) {
    uint8_t bytes[32]; fr_to_bytes(scalar, bytes);                                     /* :163 */
    for (int j = 0; j < 32; j++) for (int i = 0; i < 8; i++) out[8 * j + i] = (bytes[j] >> i) & 1;   /* :164-168 */
}

uint64_t orc_bits_count(fr_t scalar) {
    scalar = fr_reduce(&scalar);                                                       /* :174 */
    fr_t one = fr_one(), one_r = fr_reduce(&one);
    uint64_t counter = 1;
    while (fr_cmp_raw(&scalar, &one_r) > 0) { fr_divn(&scalar, 1); counter++; }        /* :176-179 */
    return counter;
}

uint64_t orc_num_bits_closest_power_of_two(fr_t scalar) {
    uint64_t num_bits = orc_bits_count(scalar);                                        /* :186 */
    fr_t closest = fr_pow_of_2(num_bits);                                              /* :187 */
    return orc_bits_count(closest);                                                    /* :188 */
}

uint64_t orc_scalar_decomposition_gadget(orc_composer *c, size_t num_bits, orc_allocated_scalar witness) {
    uint8_t scalar_bits[256]; orc_scalar_to_bits(&witness.scalar, scalar_bits);        /* :125 */
    uint64_t bit_vars[256];
    for (int i = 0; i < 256; i++) { fr_t b = fr_from_u64(scalar_bits[i]); bit_vars[i] = orc_add_input(c, &b); }   /* :128-131 */
    if (num_bits > 256) abort();                                                       /* :134 slice would panic */
    fr_t zero = fr_zero(), one = fr_one();
    orc_allocated_scalar acc; acc.var = orc_add_witness_to_circuit_description(c, &zero); acc.scalar = zero;   /* :138-141 */
    for (size_t power = 0; power < num_bits; power++) {                                /* :143-153 */
        orc_boolean_gate(c, bit_vars[power]);
        fr_t tp = two_pow((uint64_t)power);
        acc.var = orc_add(c, &tp, bit_vars[power], &one, acc.var, &zero, NULL);
        fr_t bs = fr_from_u64(scalar_bits[power]); fr_t t = fr_mul(&tp, &bs);
        acc.scalar = fr_add(&acc.scalar, &t);
    }
    return orc_maybe_equal(c, acc, witness);                                           /* :155 */
}

uint64_t orc_max_bound(orc_composer *c, fr_t max_range, orc_allocated_scalar witness, uint64_t *num_bits) {
    fr_t one = fr_one(), zero = fr_zero(), m1 = fr_neg(&one);
    max_range = fr_sub(&max_range, &one);                                              /* :87 */
    uint64_t k = orc_num_bits_closest_power_of_two(max_range);                         /* :90 */
    uint64_t b_minus_x_var = orc_add(c, &m1, witness.var, &zero, witness.var, &max_range, NULL);   /* :93-99 */
    orc_allocated_scalar v; v.var = b_minus_x_var; v.scalar = fr_sub(&max_range, &witness.scalar); /* :102-107 */
    if (num_bits) *num_bits = k;
    return orc_scalar_decomposition_gadget(c, (size_t)k, v);                           /* :110 */
}

uint64_t orc_min_bound(orc_composer *c, fr_t min_range, orc_allocated_scalar witness, uint64_t num_bits) {
    fr_t one = fr_one(), zero = fr_zero(), q_c = fr_neg(&min_range);
    uint64_t x_min_a_var = orc_add(c, &one, witness.var, &zero, witness.var, &q_c, NULL);          /* :60-66 */
    orc_allocated_scalar v; v.var = x_min_a_var; v.scalar = fr_sub(&witness.scalar, &min_range);   /* :69-74 */
    return orc_scalar_decomposition_gadget(c, (size_t)num_bits, v);                    /* :75 */
}

uint64_t orc_range_check(orc_composer *c, fr_t min_range, fr_t max_range, orc_allocated_scalar witness) {
    uint64_t k; fr_t one = fr_one(), zero = fr_zero();
    uint64_t y1 = orc_max_bound(c, max_range, witness, &k);                            /* :34 */
    uint64_t y2 = orc_min_bound(c, min_range, witness, k);                             /* :37 */
    return orc_mul(c, &one, y1, y2, &zero, NULL);                                      /* :42 */
}
