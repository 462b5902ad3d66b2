// templates.hpp -- row templates of the six gadgets (host-side, no CUDA dependency).
//
// A template is the data-independent part of what one gadget call appends to the reference composer: the order in which
// variables are allocated, and for every row its four wires and its selector values.  (The structure depends only on the
// public bit width k, never on witness values -- the property the reference's verifier-side rebuild relies on,
// /root/reference/tests/scalar_gadgets_tests.rs:43,:60.)  The builders below are written as the gadget bodies of
// /root/reference/src/range.rs and /root/reference/src/scalar.rs against a symbolic composer (`TemplateComposer`) whose
// methods carry the names and semantics of dusk-plonk's StandardComposer, so each builder can be
// read side by side with the reference function it restates.
#pragma once
#include <stdint.h>
#include <string.h>
#include <array>
#include <map>
#include <vector>
#include "bodies.cuh"

namespace pg {

enum GadgetKind : int {
    G_PREAMBLE = 0, G_ADD_INPUT, G_RANGE_CHECK, G_MAX_BOUND, G_MAYBE_EQUAL, G_IS_NON_ZERO, G_IS_NON_ZERO_PARTIAL,
    G_SELECT_ZERO, G_SELECT_ONE, G_CONSTRAIN, G_RANGE_GATE
};

// symbolic Variable: the zero variable, a local variable of this gadget instance, or an operand (a Variable that existed
// before the call: `witness.var`, `a.var`, `x`, `select`, ...)
struct WireRef { uint8_t src; uint32_t idx; };
inline WireRef W_ZERO() { WireRef w = {0, 0}; return w; }
inline WireRef W_LOCAL(uint32_t j) { WireRef w = {1, j}; return w; }
inline WireRef W_OPERAND(uint32_t e) { WireRef w = {(uint8_t)(2 + e), 0}; return w; }

// selector operand: a constant (pool index) or a per-instance parameter slot
struct SelRef { uint16_t pool; int16_t param; };
inline SelRef S_POOL(uint16_t p) { SelRef s = {p, -1}; return s; }
inline SelRef S_PARAM(int16_t p) { SelRef s = {POOL_ZERO, p}; return s; }

// gate: GATE_* (layout.h).  perm: how the row enters perm.variable_map -- PERM_LROF: add_variables_to_map(a, b, c, d) (every arithmetic
// gate method); PERM_FORL: range_gate's add_wire order Fourth, Output, Right, Left; PERM_F_ONLY: range_gate's closing gate, whose
// w_l, w_r, w_o are pushed without a map entry (their positions stay fixed points of sigma).
enum : uint8_t { PERM_LROF = 0, PERM_FORL = 1, PERM_F_ONLY = 2 };
struct RowT { WireRef w[4]; uint16_t sel[6]; uint16_t pi_sel; int16_t qc_param; int16_t pi_param; uint8_t gate; uint8_t perm; };

struct Template {
    int kind = 0; uint32_t k = 0;
    uint32_t n_vars = 0, n_fr = 0, n_planes = 0, n_params = 0, n_operands = 0;
    std::vector<uint32_t> var_loc;      // location (own table) of every local variable, in allocation order
    std::vector<RowT> rows;
    std::vector<Fr> pool;
    DecompSlots decomp[2] = {};         // filled by the range builders
    uint32_t slot_o = 0;
    int16_t param_m = -1, param_negmin = -1, param_qc = -1, param_pi = -1;
};

class TemplateComposer {
public:
    Template t;
    std::map<std::array<uint32_t, 8>, uint16_t> pool_index;
    TemplateComposer() {
        constant(fr_zero());                         // POOL_ZERO
        constant(fr_one());                          // POOL_ONE
        constant(fr_neg(fr_one()));                  // POOL_MINUS_ONE
    }
    // index of a selector constant in the pool (each distinct value once)
    uint16_t constant(const Fr& v) {
        std::array<uint32_t, 8> key;
        for (int i = 0; i < 8; i++) key[i] = v.v[i];
        auto it = pool_index.find(key);
        if (it != pool_index.end()) return it->second;
        t.pool.push_back(v);
        const uint16_t idx = (uint16_t)(t.pool.size() - 1);
        pool_index[key] = idx;
        return idx;
    }
    // composer.add_input(scalar): a new variable holding a full scalar -> one fr slot
    WireRef add_input() {
        t.var_loc.push_back(loc_make(LOC_FR, 0, t.n_fr++));
        return W_LOCAL(t.n_vars++);
    }
    // 256 x composer.add_input(BlsScalar::from(bit)) -> one bit plane (range.rs:128-131)
    uint32_t add_input_bits256() {
        const uint32_t plane = t.n_planes++, first = t.n_vars;
        for (uint32_t b = 0; b < 256; b++) t.var_loc.push_back(loc_make(LOC_BIT, 0, plane * 256 + b));
        t.n_vars += 256;
        return first;
    }
    int16_t new_param() { return (int16_t)t.n_params++; }

    void push_row(WireRef a, WireRef b, WireRef c, WireRef d, uint16_t q_m, uint16_t q_l, uint16_t q_r, uint16_t q_o, uint16_t q_4,
                  SelRef q_c, SelRef pi) {
        RowT r; r.w[0] = a; r.w[1] = b; r.w[2] = c; r.w[3] = d;
        r.sel[0] = q_m; r.sel[1] = q_l; r.sel[2] = q_r; r.sel[3] = q_o; r.sel[4] = q_4; r.sel[5] = q_c.pool;
        r.qc_param = q_c.param; r.pi_sel = pi.pool; r.pi_param = pi.param; r.gate = GATE_ARITH; r.perm = PERM_LROF;
        t.rows.push_back(r);
    }
    // a gate of range_gate: wires given in the order the quads are laid out (w_4, w_o, w_r, w_l), every arithmetic selector 0
    void push_range_row(WireRef d, WireRef o, WireRef r_, WireRef l, bool last) {
        push_row(l, r_, o, d, POOL_ZERO, POOL_ZERO, POOL_ZERO, POOL_ZERO, POOL_ZERO, S_POOL(POOL_ZERO), S_POOL(POOL_ZERO));
        t.rows.back().gate = last ? GATE_NONE : GATE_RANGE;
        t.rows.back().perm = last ? PERM_F_ONLY : PERM_FORL;
    }
    // poly_gate(a,b,c,q_m,q_l,q_r,q_o,q_c,pi): wires (a,b,c,zero_var), q_4 = 0
    void poly_gate(WireRef a, WireRef b, WireRef c, uint16_t q_m, uint16_t q_l, uint16_t q_r, uint16_t q_o, SelRef q_c, SelRef pi = S_POOL(POOL_ZERO)) {
        push_row(a, b, c, W_ZERO(), q_m, q_l, q_r, q_o, POOL_ZERO, q_c, pi);
    }
    // constrain_to_constant(a, constant, pi) = poly_gate(a,a,a, 0,1,0,0, -constant, pi); caller passes -constant
    void constrain_to_constant(WireRef a, SelRef neg_constant, SelRef pi = S_POOL(POOL_ZERO)) {
        poly_gate(a, a, a, POOL_ZERO, POOL_ONE, POOL_ZERO, POOL_ZERO, neg_constant, pi);
    }
    // add_witness_to_circuit_description(value): add_input + constrain_to_constant (never cached)
    WireRef add_witness_to_circuit_description(const Fr& value) {
        WireRef v = add_input();
        constrain_to_constant(v, S_POOL(constant(fr_neg(value))));
        return v;
    }
    // assert_equal(a,b) = poly_gate(a,b,zero, 0,1,-1,0,0)
    void assert_equal(WireRef a, WireRef b) { poly_gate(a, b, W_ZERO(), POOL_ZERO, POOL_ONE, POOL_MINUS_ONE, POOL_ZERO, S_POOL(POOL_ZERO)); }
    // add((q_l,a),(q_r,b),q_c,None): new output variable c, row (a,b,c,zero | 0,q_l,q_r,-1,0,q_c)
    WireRef add(uint16_t q_l, WireRef a, uint16_t q_r, WireRef b, SelRef q_c) {
        WireRef c = add_input();
        push_row(a, b, c, W_ZERO(), POOL_ZERO, q_l, q_r, POOL_MINUS_ONE, POOL_ZERO, q_c, S_POOL(POOL_ZERO));
        return c;
    }
    // mul(q_m,a,b,q_c,None): new output variable c, row (a,b,c,zero | q_m,0,0,-1,0,q_c)
    WireRef mul(uint16_t q_m, WireRef a, WireRef b, SelRef q_c) {
        WireRef c = add_input();
        push_row(a, b, c, W_ZERO(), q_m, POOL_ZERO, POOL_ZERO, POOL_MINUS_ONE, POOL_ZERO, q_c, S_POOL(POOL_ZERO));
        return c;
    }
    // mul_gate(a,b,c,q_m,q_o,q_c,None): row only
    void mul_gate(WireRef a, WireRef b, WireRef c, uint16_t q_m, uint16_t q_o, SelRef q_c) {
        push_row(a, b, c, W_ZERO(), q_m, POOL_ZERO, POOL_ZERO, q_o, POOL_ZERO, q_c, S_POOL(POOL_ZERO));
    }
    // boolean_gate(a): row (a,a,a,zero | 1,0,0,-1,0,0)
    void boolean_gate(WireRef a) {
        push_row(a, a, a, W_ZERO(), POOL_ONE, POOL_ZERO, POOL_ZERO, POOL_MINUS_ONE, POOL_ZERO, S_POOL(POOL_ZERO), S_POOL(POOL_ZERO));
    }
};

// ---- scalar.rs ---------------------------------------------------------------------------------------------------------
// maybe_equal(composer, a, b) -- scalar.rs:105-140.  Local variables in allocation order: u, z, y.
inline WireRef tmpl_maybe_equal(TemplateComposer& c, WireRef a, WireRef b, uint32_t* slot_u = nullptr) {
    WireRef u = c.add(POOL_ONE, a, POOL_MINUS_ONE, b, S_POOL(POOL_ZERO));      // :111-117
    if (slot_u) *slot_u = loc_payload(c.t.var_loc[u.idx]);
    WireRef z = c.add_input();                                                  // :123
    WireRef y = c.mul(POOL_MINUS_ONE, z, u, S_POOL(POOL_ONE));                  // :126
    c.mul_gate(y, u, u, POOL_ONE, POOL_ZERO, S_POOL(POOL_ZERO));                // :129-138
    return y;
}
// conditionally_select_zero -- scalar.rs:21-27
inline WireRef tmpl_select_zero(TemplateComposer& c, WireRef x, WireRef select) {
    return c.mul(POOL_ONE, x, select, S_POOL(POOL_ZERO));
}
// conditionally_select_one -- scalar.rs:36-59.  Locals: one, selector_y, one_min_selector, result.
inline WireRef tmpl_select_one(TemplateComposer& c, WireRef y, WireRef selector) {
    WireRef one = c.add_witness_to_circuit_description(fr_one());               // :41
    WireRef selector_y = c.mul(POOL_ONE, y, selector, S_POOL(POOL_ZERO));       // :43
    WireRef one_min_selector = c.add(POOL_ONE, one, POOL_MINUS_ONE, selector, S_POOL(POOL_ZERO));   // :45-50
    return c.add(POOL_ONE, selector_y, POOL_ONE, one_min_selector, S_POOL(POOL_ZERO));              // :53-58
}
// is_non_zero -- scalar.rs:63-97.  Locals: var_assigned, inv, one.  partial = the state left by the early return at :79.
inline void tmpl_is_non_zero(TemplateComposer& c, WireRef var, bool partial) {
    WireRef var_assigned = c.add_input();                                       // :69
    c.assert_equal(var, var_assigned);                                          // :71
    if (partial) return;                                                        // :79
    WireRef inv = c.add_input();                                                // :77
    WireRef one = c.add_witness_to_circuit_description(fr_one());               // :83
    c.poly_gate(var, inv, one, POOL_ONE, POOL_ZERO, POOL_ZERO, POOL_MINUS_ONE, S_POOL(POOL_ZERO));   // :84-94
}

// ---- range.rs ----------------------------------------------------------------------------------------------------------
// scalar_decomposition_gadget(composer, num_bits, witness) -- range.rs:119-158.
// Locals: 256 bit variables, A_0, A_1..A_k, then maybe_equal's u, z, y.
inline WireRef tmpl_scalar_decomposition(TemplateComposer& c, uint32_t num_bits, WireRef witness, DecompSlots* slots) {
    const uint32_t first_bit = c.add_input_bits256();                           // :128-131 (all 256 are allocated)
    slots->plane = loc_payload(c.t.var_loc[first_bit]) >> 8;
    WireRef acc = c.add_witness_to_circuit_description(fr_zero());              // :138-141
    slots->a0 = loc_payload(c.t.var_loc[acc.idx]);
    for (uint32_t power = 0; power < num_bits; power++) {                       // :143-153 (only the first num_bits, :134)
        WireRef bit = W_LOCAL(first_bit + power);
        c.boolean_gate(bit);                                                    // :144
        acc = c.add(c.constant(h_pow2[power]), bit, POOL_ONE, acc, S_POOL(POOL_ZERO));   // :146-151
    }
    WireRef y = tmpl_maybe_equal(c, acc, witness, &slots->u);                   // :155
    slots->z = slots->u + 1; slots->y = slots->u + 2;
    return y;
}
// max_bound -- range.rs:82-113.  q_c = max_range - 1 (pool constant or per-instance parameter).
inline WireRef tmpl_max_bound(TemplateComposer& c, uint32_t k, WireRef witness, SelRef max_minus_one, DecompSlots* slots) {
    WireRef b_minus_x = c.add(POOL_MINUS_ONE, witness, POOL_ZERO, witness, max_minus_one);   // :93-99
    slots->v = loc_payload(c.t.var_loc[b_minus_x.idx]);
    return tmpl_scalar_decomposition(c, k, b_minus_x, slots);                   // :110 via range_proof :21-24
}
// min_bound -- range.rs:53-76.  q_c = -min_range.
inline WireRef tmpl_min_bound(TemplateComposer& c, uint32_t k, WireRef witness, SelRef neg_min, DecompSlots* slots) {
    WireRef x_min_a = c.add(POOL_ONE, witness, POOL_ZERO, witness, neg_min);    // :60-66
    slots->v = loc_payload(c.t.var_loc[x_min_a.idx]);
    return tmpl_scalar_decomposition(c, k, x_min_a, slots);                     // :75
}
// range_check -- range.rs:27-43
inline WireRef tmpl_range_check(TemplateComposer& c, uint32_t k, WireRef witness, SelRef max_minus_one, SelRef neg_min) {
    WireRef y1 = tmpl_max_bound(c, k, witness, max_minus_one, &c.t.decomp[0]);  // :34
    WireRef y2 = tmpl_min_bound(c, k, witness, neg_min, &c.t.decomp[1]);        // :37
    WireRef o = c.mul(POOL_ONE, y1, y2, S_POOL(POOL_ZERO));                     // :42
    c.t.slot_o = loc_payload(c.t.var_loc[o.idx]);
    return o;
}

// ---- dusk-plonk's native range gate --------------------------------------------------------------------------
// StandardComposer::range_gate(witness, num_bits) [dusk-plonk 0.8 src/constraint_system/range.rs, recalled], the path
// /root/reference/src/range.rs:9-12 recommends when the bound is a power of two.  num_bits even, 2..256.
//   num_gates = ceil(num_bits / 8), num_quads = 4 * num_gates, pad = 1 + (2 * num_quads - num_bits) / 2   (1..4 zero wires first)
//   positions 0..=num_quads: `pad` times the zero variable, then the num_bits/2 accumulators (new variables, most significant
//   quad first); position i sits in gate i/4 on wire w_4, w_o, w_r, w_l for i%4 = 0, 1, 2, 3; the closing gate holds the last
//   accumulator on w_4 and zero wires elsewhere and has q_range = 0; then assert_equal(last accumulator, witness).
// Rows: num_gates + 2, variables: num_bits / 2.
inline void tmpl_range_gate(TemplateComposer& c, WireRef witness, uint32_t num_bits) {
    uint32_t num_gates = num_bits >> 3;
    if (num_bits % 8 != 0) num_gates += 1;
    const uint32_t num_quads = num_gates * 4;
    const uint32_t pad = 1 + (((num_quads << 1) - num_bits) >> 1);
    std::vector<WireRef> pos(num_quads + 1, W_ZERO());
    for (uint32_t i = pad; i <= num_quads; i++) pos[i] = c.add_input();
    for (uint32_t g = 0; g < num_gates; g++) c.push_range_row(pos[4 * g], pos[4 * g + 1], pos[4 * g + 2], pos[4 * g + 3], false);
    c.push_range_row(pos[num_quads], W_ZERO(), W_ZERO(), W_ZERO(), true);
    c.assert_equal(pos[num_quads], witness);
}

// ---- whole-call templates ------------------------------------------------------------------------------------------------
// uniform bounds: m / negmin are constants; otherwise two per-instance parameter slots
inline Template make_range_template(bool range_check, uint32_t k, bool uniform, const Fr& m, const Fr& negmin, uint32_t* result_local) {
    TemplateComposer c; c.t.kind = range_check ? G_RANGE_CHECK : G_MAX_BOUND; c.t.k = k; c.t.n_operands = 1;
    SelRef sm, sn = S_POOL(POOL_ZERO);
    if (uniform) { sm = S_POOL(c.constant(m)); if (range_check) sn = S_POOL(c.constant(negmin)); }
    else { c.t.param_m = c.new_param(); sm = S_PARAM(c.t.param_m); if (range_check) { c.t.param_negmin = c.new_param(); sn = S_PARAM(c.t.param_negmin); } }
    WireRef r = range_check ? tmpl_range_check(c, k, W_OPERAND(0), sm, sn) : tmpl_max_bound(c, k, W_OPERAND(0), sm, &c.t.decomp[0]);
    *result_local = r.idx;
    return c.t;
}
inline Template make_maybe_equal_template(uint32_t* result_local) {
    TemplateComposer c; c.t.kind = G_MAYBE_EQUAL; c.t.n_operands = 2;
    *result_local = tmpl_maybe_equal(c, W_OPERAND(0), W_OPERAND(1)).idx;
    return c.t;
}
inline Template make_is_non_zero_template(bool partial) {
    TemplateComposer c; c.t.kind = partial ? G_IS_NON_ZERO_PARTIAL : G_IS_NON_ZERO; c.t.n_operands = 1;
    tmpl_is_non_zero(c, W_OPERAND(0), partial);
    return c.t;
}
inline Template make_select_template(bool one, uint32_t* result_local) {
    TemplateComposer c; c.t.kind = one ? G_SELECT_ONE : G_SELECT_ZERO; c.t.n_operands = 2;
    *result_local = (one ? tmpl_select_one(c, W_OPERAND(0), W_OPERAND(1)) : tmpl_select_zero(c, W_OPERAND(0), W_OPERAND(1))).idx;
    return c.t;
}
inline Template make_range_gate_template(uint32_t num_bits) {
    TemplateComposer c; c.t.kind = G_RANGE_GATE; c.t.k = num_bits; c.t.n_operands = 1;
    tmpl_range_gate(c, W_OPERAND(0), num_bits);
    return c.t;
}
inline Template make_add_input_template() {
    TemplateComposer c; c.t.kind = G_ADD_INPUT; c.add_input();
    return c.t;
}
// constrain_to_constant(a, constant, pi): uniform values become constants, per-instance ones parameters
inline Template make_constrain_template(bool const_uniform, const Fr& neg_constant, bool has_pi, bool pi_uniform, const Fr& pi) {
    TemplateComposer c; c.t.kind = G_CONSTRAIN; c.t.n_operands = 1;
    SelRef qc, p = S_POOL(POOL_ZERO);
    if (const_uniform) qc = S_POOL(c.constant(neg_constant)); else { c.t.param_qc = c.new_param(); qc = S_PARAM(c.t.param_qc); }
    if (has_pi) { if (pi_uniform) p = S_POOL(c.constant(pi)); else { c.t.param_pi = c.new_param(); p = S_PARAM(c.t.param_pi); } }
    c.constrain_to_constant(W_OPERAND(0), qc, p);
    return c.t;
}
// StandardComposer::new(): zero_var = add_witness_to_circuit_description(0), then add_dummy_constraints():
// variables 6, 1, 7, -20; rows (6,7,-20,1 | 1,2,3,4,1,4) and (-20,6,7,zero | 1,1,1,1,0,127)   [SURVEY.md Appendix A.2]
inline Template make_preamble_template(std::vector<Fr>* values) {
    TemplateComposer c; c.t.kind = G_PREAMBLE;
    auto from_u64 = [](uint64_t v) { Fr r = {{(uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0}}; return fr_to_mont(r); };
    WireRef zero = c.add_witness_to_circuit_description(fr_zero());
    WireRef six = c.add_input(), one = c.add_input(), seven = c.add_input(), min_twenty = c.add_input();
    values->clear();
    values->push_back(fr_zero()); values->push_back(from_u64(6)); values->push_back(from_u64(1));
    values->push_back(from_u64(7)); values->push_back(fr_neg(from_u64(20)));
    c.push_row(six, seven, min_twenty, one, c.constant(from_u64(1)), c.constant(from_u64(2)), c.constant(from_u64(3)),
               c.constant(from_u64(4)), c.constant(from_u64(1)), S_POOL(c.constant(from_u64(4))), S_POOL(POOL_ZERO));
    c.push_row(min_twenty, six, seven, zero, POOL_ONE, POOL_ONE, POOL_ONE, POOL_ONE, POOL_ZERO, S_POOL(c.constant(from_u64(127))), S_POOL(POOL_ZERO));
    return c.t;
}

}  // namespace pg
