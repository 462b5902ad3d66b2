// shard.hpp -- splitting a mixed circuit (a list of batched gadget calls) over the GPUs of one box (host-side, no CUDA dependency).
//
// Gadget instances are independent: each call of the reference touches only the variables it allocates itself, its operands and
// the zero variable (/root/reference/src/range.rs:119-158 allocates its own A_0; /root/reference/src/scalar.rs:41, :83 their own ONE),
// so the instances of every call can be cut into contiguous ranges, one per rank, and each rank appends its ranges to a composer
// of its own.  What has to be computed is where a rank's rows and Variables sit in the SEQUENTIAL composer of the whole circuit
// (the thing the reference would have built): the sequential composer appends call after call, so
//       row id  of (call k, instance i, local row r)      = 3 + sum_{j<k} n_j*rows_j + i*rows_k + r
//       Variable of (call k, instance i, local variable v) = 5 + sum_{j<k} n_j*vars_j + i*vars_k + v
// (3 rows / 5 variables: StandardComposer::new()).  pg_shard_plan returns, per rank and call, the instance range
// and the row / Variable index of its first instance -- prefix sums over the per-call row and Variable counts.
#pragma once
#include <stdint.h>
#include "../../include/pg_b200.h"
#include "templates.hpp"

namespace pg {

// rows and variables ONE instance of a gadget appends: k = num_bits of the range gadgets
inline bool op_shape(uint32_t gadget, uint32_t k, uint64_t* rows, uint64_t* vars) {
    uint64_t r = 0, v = 0;
    switch (gadget) {
        case PG_OP_ADD_INPUT: r = 0; v = 1; break;                                       // allocated_scalar.rs:27-30
        case PG_OP_RANGE_CHECK: if (k < 2 || k > 256) return false; r = 4ull * k + 11; v = 2ull * k + 523; break;   // range.rs:27-43
        case PG_OP_MAX_BOUND: if (k < 2 || k > 256) return false; r = 2ull * k + 5; v = (uint64_t)k + 261; break;   // range.rs:82-113
        case PG_OP_MAYBE_EQUAL: r = 3; v = 3; break;                                     // scalar.rs:105-140
        case PG_OP_IS_NON_ZERO: r = 3; v = 3; break;                                     // scalar.rs:63-97 (PG_NZ_UNIFORM numbering)
        case PG_OP_SELECT_ZERO: r = 1; v = 1; break;                                     // scalar.rs:21-27
        case PG_OP_SELECT_ONE: r = 4; v = 4; break;                                      // scalar.rs:36-59
        case PG_OP_CONSTRAIN: r = 1; v = 0; break;                                       // constrain_to_constant [dusk-plonk]
        case PG_OP_RANGE_GATE: if (k < 2 || k > 256 || (k & 1)) return false; r = (k + 7) / 8 + 2; v = k / 2; break;   // range_gate [dusk-plonk]
        default: return false;
    }
    if (rows) *rows = r;
    if (vars) *vars = v;
    return true;
}

constexpr uint64_t FRESH_ROWS = 3, FRESH_VARS = 5;     // StandardComposer::new()

// pg_template_get: the rows ONE instance of a gadget appends, in the reference's order -- what "emits the same selector rows" means,
// without materialising a batch.  Wire references: 0 = the zero variable, 1 + j = the j-th Variable the instance allocates (Variable id
// first + j), -1 - e = operand e (the Variable(s) passed in: witness / a, b / x, select / var).  Selectors q_m q_l q_r q_o q_4 q_c as values;
// entries that depend on per-instance bounds / constants are the ones of the bounds passed here.  gate: 0 arithmetic (q_arith = 1),
// 1 range widget (q_range = 1), 2 neither.  Buffers may be null (sizes only).
inline int template_get(uint32_t gadget, uint32_t num_bits, const pg_fr* mn, const pg_fr* mx, uint64_t* n_rows, uint64_t* n_vars,
                        int64_t* w_ref, pg_fr* sel, uint32_t* gate) {
    auto fr_of = [](const pg_fr* x) { Fr r = fr_zero(); if (x) for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)x->l[i]; r.v[2 * i + 1] = (uint32_t)(x->l[i] >> 32); } return r; };
    Template t; uint32_t result_local = 0;
    if (fr_is_zero(h_pow2[0])) {                               // the 2^i table of the range templates (filled by the first ctx otherwise)
        h_pow2[0] = fr_one();
        for (int i = 1; i < 256; i++) h_pow2[i] = fr_add(h_pow2[i - 1], h_pow2[i - 1]);
    }
    switch (gadget) {
        case PG_OP_ADD_INPUT: t = make_add_input_template(); break;
        case PG_OP_RANGE_CHECK: case PG_OP_MAX_BOUND: {
            if (!mx || (gadget == PG_OP_RANGE_CHECK && !mn)) return PG_ERR_ARG;
            const Fr m = fr_sub(fr_of(mx), fr_one());
            const uint32_t k = num_bits_from_canonical(fr_from_mont(m));
            if (num_bits && num_bits != k) return PG_ERR_ARG;                      // num_bits is implied by the bound (range.rs:87-90)
            t = make_range_template(gadget == PG_OP_RANGE_CHECK, k, true, m, fr_neg(fr_of(mn)), &result_local);
        } break;
        case PG_OP_MAYBE_EQUAL: t = make_maybe_equal_template(&result_local); break;
        case PG_OP_IS_NON_ZERO: t = make_is_non_zero_template(false); break;
        case PG_OP_SELECT_ZERO: t = make_select_template(false, &result_local); break;
        case PG_OP_SELECT_ONE: t = make_select_template(true, &result_local); break;
        case PG_OP_CONSTRAIN: t = make_constrain_template(true, fr_neg(fr_of(mx)), mn != nullptr, true, fr_of(mn)); break;   // mx: the constant, mn: the PI (or null)
        case PG_OP_RANGE_GATE: if (num_bits < 2 || num_bits > 256 || (num_bits & 1)) return PG_ERR_ARG; t = make_range_gate_template(num_bits); break;
        default: return PG_ERR_ARG;
    }
    const uint64_t R = t.rows.size();
    if (n_rows) *n_rows = R;
    if (n_vars) *n_vars = t.n_vars;
    for (uint64_t r = 0; r < R; r++) {
        const RowT& row = t.rows[r];
        for (int w = 0; w < 4 && w_ref; w++) {
            const WireRef& wr = row.w[w];
            w_ref[(uint64_t)w * R + r] = wr.src == 0 ? 0 : wr.src == 1 ? 1 + (int64_t)wr.idx : -1 - (int64_t)(wr.src - 2);
        }
        for (int k = 0; k < 6 && sel; k++) {
            const Fr v = t.pool[row.sel[k]];
            pg_fr& o = sel[(uint64_t)k * R + r];
            for (int i = 0; i < 4; i++) o.l[i] = (uint64_t)v.v[2 * i] | ((uint64_t)v.v[2 * i + 1] << 32);
        }
        if (gate) gate[r] = row.gate;
    }
    return PG_OK;
}

// out[rank * n_ops + k] = what `rank` runs of call k.  Calls with the same `group` share one instance index space (a column and the
// gadgets applied to it: instance i of each of them must live on the same rank) and must have the same n; groups are numbered in
// order of first appearance and a group's calls need not be adjacent.
//   PG_SHARD_EVEN: every group is cut into `world` equal instance ranges (floor(n*r/G) .. floor(n*(r+1)/G)).
//   PG_SHARD_ROWS: the groups, in order, are laid end to end weighted by the rows one instance of the group appends, and that
//                  line is cut into `world` equal parts at instance boundaries: a rank gets a contiguous run of the circuit with
//                  ~1/G of the rows (the gate check costs the same for every row).  Groups without rows are cut evenly.
inline int shard_plan(const pg_op* ops, uint64_t n_ops, uint32_t world, int policy, pg_op_shard* out) {
    if (!world || (n_ops && (!ops || !out)) || (policy != PG_SHARD_EVEN && policy != PG_SHARD_ROWS)) return PG_ERR_ARG;
    // global bases: prefix sums over the calls
    uint64_t row = FRESH_ROWS, var = FRESH_VARS;
    for (uint64_t k = 0; k < n_ops; k++) {
        uint64_t r, v;
        if (!op_shape(ops[k].gadget, ops[k].num_bits, &r, &v)) return PG_ERR_ARG;
        for (uint64_t j = 0; j < k; j++) if (ops[j].group == ops[k].group && ops[j].n != ops[k].n) return PG_ERR_ARG;
        for (uint32_t g = 0; g < world; g++) { pg_op_shard& s = out[(uint64_t)g * n_ops + k]; s.row_base = row; s.var_base = var; s.inst_lo = s.inst_hi = 0; }
        row += ops[k].n * r; var += ops[k].n * v;
    }
    // groups in order of first appearance: instance count and rows per instance
    unsigned __int128 total_w = 0;
    for (uint64_t k = 0; k < n_ops; k++) {
        bool first = true;
        for (uint64_t j = 0; j < k; j++) if (ops[j].group == ops[k].group) { first = false; break; }
        if (!first) continue;
        uint64_t w = 0;
        for (uint64_t j = k; j < n_ops; j++) if (ops[j].group == ops[k].group) { uint64_t r; op_shape(ops[j].gadget, ops[j].num_bits, &r, nullptr); w += r; }
        total_w += (unsigned __int128)w * ops[k].n;
    }
    unsigned __int128 prefix = 0;
    for (uint64_t k = 0; k < n_ops; k++) {
        bool first = true;
        for (uint64_t j = 0; j < k; j++) if (ops[j].group == ops[k].group) { first = false; break; }
        if (!first) continue;
        const uint64_t n = ops[k].n;
        uint64_t w = 0;
        for (uint64_t j = k; j < n_ops; j++) if (ops[j].group == ops[k].group) { uint64_t r; op_shape(ops[j].gadget, ops[j].num_bits, &r, nullptr); w += r; }
        for (uint32_t g = 0; g < world; g++) {
            uint64_t lo, hi;
            if (policy == PG_SHARD_EVEN || w == 0 || total_w == 0) {
                lo = (uint64_t)((unsigned __int128)n * g / world); hi = (uint64_t)((unsigned __int128)n * (g + 1) / world);
            } else {
                // instance i starts at prefix + i*w on the weighted line; it belongs to the rank whose part [T_g, T_{g+1}) holds that point
                auto cut = [&](uint32_t r) -> uint64_t {
                    const unsigned __int128 T = total_w * r / world;
                    if (r >= world) return n;
                    if (T <= prefix) return 0;
                    const unsigned __int128 c = (T - prefix + w - 1) / w;
                    return c > n ? n : (uint64_t)c;
                };
                lo = cut(g); hi = cut(g + 1);
            }
            for (uint64_t j = k; j < n_ops; j++) {
                if (ops[j].group != ops[k].group) continue;
                uint64_t r, v; op_shape(ops[j].gadget, ops[j].num_bits, &r, &v);
                pg_op_shard& s = out[(uint64_t)g * n_ops + j];
                s.inst_lo = lo; s.inst_hi = hi; s.row_base += lo * r; s.var_base += lo * v;
            }
        }
        prefix += (unsigned __int128)w * n;
    }
    return PG_OK;
}

}  // namespace pg
