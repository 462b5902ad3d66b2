// kernels.cuh -- sm_100a kernels: thin __global__ wrappers around the per-instance bodies (bodies.cuh), the block-wide
// Montgomery batch inversion, the gate-check kernel and the integer-multiply roofline micro-benchmarks.
#pragma once
#include <cuda_runtime.h>
#include "bodies.cuh"

namespace pg {

constexpr int BLOCK = 256;            // threads per block of the per-instance kernels (8 warps)
constexpr int NWARPS = BLOCK / 32;

// ---------------------------------------------------------------------------------------------------- simple wrapper
template <class Body>
__global__ void __launch_bounds__(BLOCK) k_simple(const typename Body::Args a) {
    const uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i < a.n) Body::run(a, i);
}

// ---------------------------------------------------------------------------------------------------- batch inversion
__device__ __forceinline__ Fr shfl_xor_fr(const Fr& a, int mask) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = __shfl_xor_sync(0xffffffffu, a.v[k], mask);
    return r;
}

// Montgomery's trick across a whole thread block, zero-aware.
//   in : v[e] arbitrary field elements (zeros allowed), E per thread, every thread of the block must call
//   out: v[e] = v[e]^-1, or 0 where v[e] was 0          (== BlsScalar::invert().unwrap_or(zero), scalar.rs:122)
// Products are combined with an xor-butterfly over warp shuffles (each lane keeps the sibling product of every level),
// warp totals are staged in shared memory and combined by warp 0 with the same butterfly; ONE Fermat inversion is
// executed per block; inverses are pushed back down the two butterflies with one multiplication per level.
template <int E>
__device__ __forceinline__ void block_batch_invert(Fr (&v)[E], Fr* smem /* NWARPS entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Fr one = fr_one();
    bool nz[E]; Fr pre[E];                       // pre[e] = product of the (zero-patched) elements 0..e of this thread
#pragma unroll
    for (int e = 0; e < E; e++) {
        nz[e] = !fr_is_zero(v[e]);
        if (!nz[e]) v[e] = one;
        pre[e] = e == 0 ? v[0] : fr_mul(pre[e - 1], v[e]);
    }
    // up-sweep inside the warp
    Fr g = pre[E - 1], sib[5];
#pragma unroll
    for (int l = 0; l < 5; l++) { sib[l] = shfl_xor_fr(g, 1 << l); g = fr_mul(g, sib[l]); }
    if (lane == 0) smem[warp] = g;
    __syncthreads();
    if (warp == 0) {
        Fr h = lane < NWARPS ? smem[lane] : one, hs[3];
#pragma unroll
        for (int l = 0; l < 3; l++) { hs[l] = shfl_xor_fr(h, 1 << l); h = fr_mul(h, hs[l]); }   // NWARPS == 8 -> 3 levels
        Fr ih = fr_inv_fermat(h);                                                               // the block's only inversion
#pragma unroll
        for (int l = 2; l >= 0; l--) ih = fr_mul(ih, hs[l]);                                    // inverse of smem[lane]
        if (lane < NWARPS) smem[lane] = ih;
    }
    __syncthreads();
    Fr ig = smem[warp];                                                                         // inverse of the warp total
#pragma unroll
    for (int l = 4; l >= 0; l--) ig = fr_mul(ig, sib[l]);                                       // inverse of this thread's product
#pragma unroll
    for (int e = E - 1; e >= 1; e--) {
        const Fr inv_e = fr_mul(ig, pre[e - 1]);
        ig = fr_mul(ig, v[e]);
        v[e] = nz[e] ? inv_e : fr_zero();
    }
    v[0] = nz[0] ? ig : fr_zero();
}
static_assert(NWARPS == 8, "block_batch_invert's cross-warp butterfly is written for 8 warps");

template <class Body>
__global__ void __launch_bounds__(BLOCK) k_inv(const typename Body::Args a) {
    __shared__ Fr smem[NWARPS];
    const uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool active = i < a.n;
    typename Body::State st;
    Fr v[Body::E];
    if (active) Body::pre(a, i, st, v);
    else {
#pragma unroll
        for (int e = 0; e < Body::E; e++) v[e] = fr_one();
    }
    block_batch_invert<Body::E>(v, smem);
    if (active) Body::post(a, i, st, v);
}

// ---------------------------------------------------------------------------------------------------- gate check
struct SmemPool {
    const uint32_t* p;
    __device__ __forceinline__ Fr operator()(uint32_t idx) const {
        const uint4 lo = *reinterpret_cast<const uint4*>(p + 8 * idx), hi = *reinterpret_cast<const uint4*>(p + 8 * idx + 4);
        Fr r = {{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
        return r;
    }
};

// One instance per thread; the row template is walked by all threads in lock step (selector pool in shared memory,
// template rows read with warp-uniform addresses), wire values come from the SoA variable table with coalesced loads.
__global__ void __launch_bounds__(BLOCK, 2) k_check(const CheckArgs a) {
    extern __shared__ __align__(16) uint32_t s_pool[];
    for (uint32_t t = threadIdx.x; t < a.n_pool * 8; t += BLOCK) s_pool[t] = a.pool[t];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    unsigned long long first_bad = ~0ull;
    uint32_t bad = 0;
    if (i < a.n_inst) { SmemPool pool = {s_pool}; bad = CheckBody::run(a, pool, i, first_bad); }
    // warp-level reduction, then one atomic per warp that saw a violation
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, first_bad, o);
        first_bad = other < first_bad ? other : first_bad;
    }
    if ((threadIdx.x & 31) == 0 && bad) {
        atomicAdd(a.counters + CNT_UNSAT, (unsigned long long)bad);
        atomicMin(a.counters + CNT_FIRST_BAD, first_bad);
    }
}

__global__ void __launch_bounds__(BLOCK) k_check_rows(const CheckRowsBody::Args a) {
    const uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    uint32_t bad = i < a.n ? CheckRowsBody::run(a, i) : 0u;
    unsigned long long first_bad = bad ? (unsigned long long)i : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, first_bad, o);
        first_bad = other < first_bad ? other : first_bad;
    }
    if ((threadIdx.x & 31) == 0 && bad) {
        atomicAdd(a.counters + CNT_UNSAT, (unsigned long long)bad);
        atomicMin(a.counters + CNT_FIRST_BAD, first_bad);
    }
}

// ---------------------------------------------------------------------------------------------------- IMAD roofline
// 8 independent accumulator chains per thread so that the multiplier pipe, not the 4-cycle dependent-issue latency, is
// the limit.  `iters` x 8 multiply-accumulates per thread.
__global__ void __launch_bounds__(BLOCK) k_imad_wide(uint64_t* out, uint32_t x, uint32_t y, int iters) {
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = (uint64_t)threadIdx.x * (k + 1) + blockIdx.x;
    uint32_t a = x + threadIdx.x, b = y | 1u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
        a += 0x9e3779b9u;
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[(uint64_t)blockIdx.x * BLOCK + threadIdx.x] = s;
}
__global__ void __launch_bounds__(BLOCK) k_imad_lo(uint32_t* out, uint32_t x, uint32_t y, int iters) {
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x * (k + 1) + blockIdx.x;
    uint32_t a = x + threadIdx.x, b = y | 1u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a), "r"(b));
        a += 0x9e3779b9u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[(uint64_t)blockIdx.x * BLOCK + threadIdx.x] = s;
}

}  // namespace pg
