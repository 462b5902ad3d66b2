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
