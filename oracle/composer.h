/*
 * oracle/composer.h -- CPU restatement of the arithmetic-row subset of dusk-plonk 0.8's StandardComposer.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/fr.h).  "parity unpinned" at row level: dusk-plonk's sources are not
 * under /root/reference (dependency `dusk-plonk = "0.8"`, /root/reference/Cargo.toml:20); the semantics are
 * restated from SURVEY.md Appendix A.2 and anchored on the reference's call sites:
 *   add_input   /root/reference/src/allocated_scalar.rs:28, range.rs:130, scalar.rs:69,:77,:123
 *   add         range.rs:65,:98,:151; scalar.rs:45,:53,:116
 *   mul         range.rs:42; scalar.rs:26,:43,:126        mul_gate     scalar.rs:137
 *   poly_gate   scalar.rs:84                             boolean_gate range.rs:144
 *   assert_equal scalar.rs:71                            add_witness_to_circuit_description range.rs:139; scalar.rs:41,:83
 *   constrain_to_constant  /root/reference/tests/range_gadgets_tests.rs:26,:43; tests/scalar_gadgets_tests.rs:30,:78,:135
 *
 * The cost structure of the original is kept on purpose (this is also the timed CPU baseline): a hash map from
 * Variable to value that is consulted by add/mul, eleven selector columns and four wire columns pushed per row,
 * a Variable -> wire-position map (the permutation argument's variable_map) updated on every row.
 */
#ifndef ORACLE_COMPOSER_H
#define ORACLE_COMPOSER_H

#include "fr.h"
#include <stddef.h>

enum { ORC_QM, ORC_QL, ORC_QR, ORC_QO, ORC_Q4, ORC_QC, ORC_QARITH, ORC_QRANGE, ORC_QLOGIC, ORC_QFIXED, ORC_QVAR, ORC_NSEL };

typedef struct { fr_t *p; size_t len, cap; } orc_frvec;
typedef struct { uint64_t *p; size_t len, cap; } orc_u64vec;

typedef struct orc_composer {
    size_t n;                         /* rows */
    orc_frvec sel[ORC_NSEL];          /* q_m q_l q_r q_o q_4 q_c q_arith q_range q_logic q_fixed_group_add q_variable_group_add */
    orc_u64vec w[4];                  /* w_l w_r w_o w_4 : Variable indices */
    /* sparse public inputs, kept sorted by row (BTreeMap<usize,BlsScalar>) */
    orc_u64vec pi_pos; orc_frvec pi_val;
    uint64_t zero_var;
    /* variables: HashMap<Variable,BlsScalar>  (open addressing, keys are Variable indices) */
    uint64_t *map_key; fr_t *map_val; size_t map_cap, map_len;
    uint64_t n_vars;
    /* perm.variable_map: Variable -> list of (wire, row); singly linked lists in arrays */
    orc_u64vec perm_head;             /* per variable: index+1 of the newest entry, 0 = none */
    orc_u64vec perm_next;             /* per entry */
    orc_u64vec perm_data;             /* per entry: row*4 + wire */
} orc_composer;

orc_composer *orc_composer_new(void);          /* StandardComposer::new(): zero var + 2 dummy rows -> n=3, 5 variables */
void orc_composer_free(orc_composer *c);

uint64_t orc_add_input(orc_composer *c, const fr_t *s);
uint64_t orc_add_witness_to_circuit_description(orc_composer *c, const fr_t *v);
void orc_poly_gate(orc_composer *c, uint64_t a, uint64_t b, uint64_t o, const fr_t *q_m, const fr_t *q_l,
                   const fr_t *q_r, const fr_t *q_o, const fr_t *q_c, const fr_t *pi /* NULL = None */);
void orc_constrain_to_constant(orc_composer *c, uint64_t a, const fr_t *constant, const fr_t *pi);
void orc_assert_equal(orc_composer *c, uint64_t a, uint64_t b);
uint64_t orc_add(orc_composer *c, const fr_t *q_l, uint64_t a, const fr_t *q_r, uint64_t b, const fr_t *q_c, const fr_t *pi);
uint64_t orc_mul(orc_composer *c, const fr_t *q_m, uint64_t a, uint64_t b, const fr_t *q_c, const fr_t *pi);
void orc_mul_gate(orc_composer *c, uint64_t a, uint64_t b, uint64_t o, const fr_t *q_m, const fr_t *q_o, const fr_t *q_c, const fr_t *pi);
void orc_boolean_gate(orc_composer *c, uint64_t a);

/* StandardComposer::range_gate(witness, num_bits) [dusk-plonk 0.8 src/constraint_system/range.rs, recalled; recommended by the
 * reference for power-of-two bounds, /root/reference/src/range.rs:9-12]: quad accumulators, four per gate (w_4, w_o, w_r, w_l),
 * q_range = 1 on all used gates but the last, assert_equal(last accumulator, witness).  num_bits must be even, 2..256. */
void orc_range_gate(orc_composer *c, uint64_t witness, uint64_t num_bits);

/* variables[var] */
fr_t orc_value_of(const orc_composer *c, uint64_t var);
/* arithmetic + range part of check_circuit_satisfied: number of rows with q_arith*(...) + q_range*(sum of four deltas) != 0;
 * first such row or (uint64_t)-1 */
uint64_t orc_check(const orc_composer *c, uint64_t *first_bad);
/* construct_dense_pi_vec(): out has c->n entries */
void orc_dense_pi(const orc_composer *c, fr_t *out);

#endif
