/*
 * oracle/gadgets.h -- CPU restatement of plonk_gadgets' six gadgets.  TEST INFRASTRUCTURE ONLY (see fr.h).
 * Follows /root/reference/src/range.rs:21-189 and /root/reference/src/scalar.rs:21-140 line by line for
 * allocation order, selector values and host-side mirror arithmetic.
 */
#ifndef ORACLE_GADGETS_H
#define ORACLE_GADGETS_H

#include "composer.h"

/* /root/reference/src/allocated_scalar.rs:17-23 */
typedef struct { uint64_t var; fr_t scalar; } orc_allocated_scalar;

/* /root/reference/src/errors.rs:13-18 */
enum { ORC_OK = 0, ORC_ERR_NON_EXISTING_INVERSE = 1 };

/* Cost model of `BlsScalar::from(2).pow(&[power,0,0,0])` at range.rs:146:
 *   ORC_FAITHFUL -- recompute it with the crate's fixed 256-step ladder for every bit (what the reference does);
 *   ORC_FAST     -- look 2^i up in a table ("optimised CPU" line of the baseline).  Results are identical. */
enum { ORC_FAITHFUL = 0, ORC_FAST = 1 };
void orc_set_mode(int mode);

orc_allocated_scalar orc_allocate(orc_composer *c, const fr_t *scalar);                       /* allocated_scalar.rs:27-30 */
uint64_t orc_bits_count(fr_t scalar);                                                           /* range.rs:173-181 */
uint64_t orc_num_bits_closest_power_of_two(fr_t scalar);                                        /* range.rs:185-189 */
void orc_scalar_to_bits(const fr_t *scalar, uint8_t out[256]);                                  /* range.rs:161-170 */
uint64_t orc_scalar_decomposition_gadget(orc_composer *c, size_t num_bits, orc_allocated_scalar witness);   /* range.rs:119-158 */
uint64_t orc_max_bound(orc_composer *c, fr_t max_range, orc_allocated_scalar witness, uint64_t *num_bits);  /* range.rs:82-113 */
uint64_t orc_min_bound(orc_composer *c, fr_t min_range, orc_allocated_scalar witness, uint64_t num_bits);   /* range.rs:53-76 */
uint64_t orc_range_check(orc_composer *c, fr_t min_range, fr_t max_range, orc_allocated_scalar witness);    /* range.rs:27-43 */
uint64_t orc_conditionally_select_zero(orc_composer *c, uint64_t x, uint64_t select);           /* scalar.rs:21-27 */
uint64_t orc_conditionally_select_one(orc_composer *c, uint64_t y, uint64_t selector);          /* scalar.rs:36-59 */
int orc_is_non_zero(orc_composer *c, uint64_t var, fr_t value_assigned);                        /* scalar.rs:63-97 */
uint64_t orc_maybe_equal(orc_composer *c, orc_allocated_scalar a, orc_allocated_scalar b);      /* scalar.rs:105-140 */

#endif
