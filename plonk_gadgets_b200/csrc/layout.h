// layout.h -- data layout of the device-resident composer (shared by host logic and kernels).
//
// The reference composer (dusk-plonk StandardComposer) keeps, per row, four `Variable`s and eleven selector scalars, and a
// hash map Variable -> BlsScalar.  For n independent instances of one gadget that is n copies of the same row *structure*:
// only the variable values (and, with per-instance bounds, two q_c entries) differ.  The engine therefore stores a
// SEGMENT per batched call:
//
//   * a row TEMPLATE (rows x {4 wire references, 6 selector pool indices, optional per-instance q_c / PI parameter});
//   * a selector POOL (the distinct Fr constants of the template: 0, 1, -1, 2^i, bounds ...);
//   * the per-instance VARIABLE TABLE in structure-of-arrays form, sized n_alloc instances:
//       fr   [slot][half][n_alloc]  uint4   -- 32-byte Montgomery scalars split in two 16-byte halves, so that a warp
//                                             reading/writing slot s touches 512 contiguous bytes per half (128-bit
//                                             coalesced accesses);
//       bits [plane][word][n_alloc] u32    -- the 256 bit-variables of one decomposition packed as the canonical
//                                             little-endian 256-bit integer (bit b of the plane = Variable B_b);
//       param[slot][half][n_alloc]  uint4   -- per-instance selector/public-input values (q_c overrides, PI).
//
// Variable numbering is the reference's: Variable id of local variable j of instance i = base_var + i*n_vars + j;
// row id of local row r = base_row + i*n_rows + r (instances are appended one after another by the sequential loop).
#pragma once
#include <stdint.h>
#include "fr.cuh"

namespace pg {

// ---- value locations ------------------------------------------------------------------------------------------------
// 32-bit code: [31:30] kind, [29:27] table (0 = this segment, 1..4 = operand column e-1), [26:0] payload
enum : uint32_t { LOC_ZERO = 0u, LOC_FR = 1u, LOC_BIT = 2u };
PG_HD uint32_t loc_make(uint32_t kind, uint32_t tab, uint32_t payload) { return (kind << 30) | (tab << 27) | payload; }
PG_HD uint32_t loc_kind(uint32_t l) { return l >> 30; }
PG_HD uint32_t loc_tab(uint32_t l) { return (l >> 27) & 7u; }
PG_HD uint32_t loc_payload(uint32_t l) { return l & 0x07ffffffu; }
PG_HD uint32_t loc_with_tab(uint32_t l, uint32_t tab) { return (l & ~(7u << 27)) | (tab << 27); }

constexpr int MAX_TABS = 5;   // own table + up to 4 operand columns

struct DevTab {
    const uint4* fr;          // pre-offset by the operand's first instance
    const uint32_t* bits;
    uint64_t stride;          // n_alloc of the owning segment
    uint64_t var_base;        // Variable id of (instance 0 of this view, local variable 0)
    uint64_t var_stride;      // n_vars of the owning segment
};

struct DevRow {               // 64 bytes
    uint32_t loc[4];          // where the values of w_l, w_r, w_o, w_4 live
    uint32_t var[4];          // local Variable index inside the owning table's segment (unused for LOC_ZERO)
    uint16_t sel[6];          // pool indices: q_m q_l q_r q_o q_4 q_c
    uint16_t pi_sel;          // pool index of a uniform public input (0 = the constant zero = no PI)
    int16_t qc_param;         // >= 0: q_c is param slot qc_param of the instance (sel[5] ignored)
    int16_t pi_param;         // >= 0: PI is param slot pi_param of the instance
    uint16_t pad[7];
};
static_assert(sizeof(DevRow) == 64, "DevRow must stay 64 bytes");

// reserved pool entries
enum : uint16_t { POOL_ZERO = 0, POOL_ONE = 1, POOL_MINUS_ONE = 2 };

// ---- table accessors ------------------------------------------------------------------------------------------------
PG_HD Fr tab_load_fr(const uint4* base, uint64_t stride, uint32_t slot, uint64_t i) {
    // (base + i) first: the per-thread part is loop-invariant and the slot offset is warp-uniform (uniform datapath), so the
    // address costs adder instructions instead of a 64-bit IMAD on the multiplier pipe
    const uint4* p = base + i;
    const uint4 lo = p[(uint64_t)(2 * slot) * stride];
    const uint4 hi = p[(uint64_t)(2 * slot + 1) * stride];
    Fr r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
PG_HD void tab_store_fr(uint4* base, uint64_t stride, uint32_t slot, uint64_t i, const Fr& v) {
    uint4* p = base + i;
    p[(uint64_t)(2 * slot) * stride] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    p[(uint64_t)(2 * slot + 1) * stride] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
PG_HD uint32_t tab_load_bit(const uint32_t* bits, uint64_t stride, uint32_t plane_bit, uint64_t i) {
    const uint32_t plane = plane_bit >> 8, bit = plane_bit & 255u;
    const uint32_t w = (bits + i)[(uint64_t)(plane * 8 + (bit >> 5)) * stride];
    return (w >> (bit & 31u)) & 1u;
}
// value of a located variable for instance i
PG_HD Fr loc_load(const DevTab* tabs, uint32_t loc, uint64_t i) {
    const uint32_t kind = loc_kind(loc);
    if (kind == LOC_ZERO) return fr_zero();
    const DevTab& t = tabs[loc_tab(loc)];
    if (kind == LOC_FR) return tab_load_fr(t.fr, t.stride, loc_payload(loc), i);
    const uint32_t b = tab_load_bit(t.bits, t.stride, loc_payload(loc), i);
    const Fr one = fr_one();
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = b ? one.v[k] : 0u;
    return r;
}
// request the cache lines a later loc_load(tabs, loc, i) will touch (no destination registers)
PG_HD void loc_prefetch(const DevTab* tabs, uint32_t loc, uint64_t i) {
#if defined(__CUDA_ARCH__)
    const uint32_t kind = loc_kind(loc);
    if (kind == LOC_ZERO) return;
    const DevTab& t = tabs[loc_tab(loc)];
    if (kind == LOC_FR) {
        const uint32_t slot = loc_payload(loc);
        const uint4* p = t.fr + i;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (uint64_t)(2 * slot) * t.stride));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (uint64_t)(2 * slot + 1) * t.stride));
    } else {
        const uint32_t pb = loc_payload(loc);
        asm volatile("prefetch.global.L1 [%0];" ::"l"((t.bits + i) + (uint64_t)((pb >> 8) * 8 + ((pb & 255u) >> 5)) * t.stride));
    }
#else
    (void)tabs; (void)loc; (void)i;
#endif
}
// AoS scalar (caller memory: BlsScalar[n]) access
PG_HD Fr aos_load(const uint4* p, uint64_t i) {
    const uint4 lo = p[2 * i], hi = p[2 * i + 1];
    Fr r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
    return r;
}
PG_HD void aos_store(uint4* p, uint64_t i, const Fr& v) {
    p[2 * i] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    p[2 * i + 1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
PG_HD Fr pool_load(const uint32_t* pool, uint32_t idx) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = pool[8 * idx + k];
    return r;
}

}  // namespace pg
