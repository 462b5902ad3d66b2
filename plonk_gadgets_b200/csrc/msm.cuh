// msm.cuh -- multi-scalar multiplication in G1: sum_i scalar_i * point_i  (KZG commitment of a coefficient vector against
// the SRS powers; SURVEY.md section 8f item 2, second half; [DEP] dusk-bls12_381 `msm_variable_base`, dusk-plonk 0.8
// `CommitKey::commit`, reached from /root/reference/tests/range_gadgets_tests.rs:90-91 through prover.prove).
//
// The reference's CPU routine is a serial Pippenger; the sum it returns is a group element, independent of the schedule.  Here:
//   1. k_simple<MsmDigitsBody>    scalars out of Montgomery form, cut into W = ceil(255/c) unsigned c-bit digits;
//                                 one (key = window*2^c + digit, value = point index) pair per scalar and window;
//   2. sort_pairs                 radix sort of the W*n pairs by key (CUB on the device): every bucket becomes a contiguous run;
//   3. k_simple<MsmBoundsBody>    start and length of every bucket's run (binary search); a second, small radix sort orders the
//                                 bucket ids by decreasing length;
//      k_simple<MsmBucketBody>    runs are cut into equal parts of at most 512 entries, numbered in that order -- the lanes of a warp
//                                 get parts of (almost) equal length, and no thread walks a long run however skewed the
//                                 scalars are: mixed additions of a part's points into an XYZZ accumulator (W*n in total);
//                                 MsmGroupSumBody / MsmBucketFinishBody add a bucket's parts (two more levels);
//   4. k_simple<MsmChunkBody>     per window, chunks of L consecutive buckets: running-sum trick gives sum (d - lo) * B_d,
//                                 plus lo * (sum B_d) by double-and-add on the small factor lo;
//   5. k_simple<MsmSumBody>       tree of 16-way sums of the chunk results down to one point per window;
//   6. k_simple<MsmFinalBody>     Horner over the windows (c doublings each) and the conversion to affine.
// c is chosen from n (MsmPlan).  All group formulas are complete (g1.cuh).
#pragma once
#include <stdlib.h>
#include "g1.cuh"
#include "layout.h"

namespace pg {

struct MsmPlan { uint32_t c, n_windows, chunk; };          // chunk = L, a power of two <= 2^c
// Window widths are taken from {3, 5, 8, 15, 16}: for these the top window (255 - c*(W-1) bits) is as wide as the others or
// one bit short, so no bucket is much fuller than average (with c = 12 the top window would have 3 bits: 8 buckets of n/8
// points each, summed by 8 threads).  Buckets per window grow with n so that bucket sums (W*n additions) and bucket
// reduction (2 * W * 2^c additions) stay balanced.
inline MsmPlan msm_plan(uint64_t n) {
    uint32_t log_n = 0;
    while ((1ull << (log_n + 1)) <= n) log_n++;
    MsmPlan p;
    p.c = log_n <= 6 ? 3 : log_n <= 9 ? 5 : log_n <= 15 ? 8 : log_n <= 22 ? 15 : 16;      // measured; 17 / 20 bits are slower even at 2^22 / 2^25
    if (const char* e = getenv("PG_MSM_C")) { const int c = atoi(e); if (c >= 2 && c <= 20) p.c = (uint32_t)c; }   // tuning runs only
    p.n_windows = (255 + p.c - 1) / p.c;
    p.chunk = p.c >= 8 ? 32u : (1u << p.c);
    return p;
}

// digit w of the canonical scalar
PG_HD uint32_t msm_digit(const Fr& canon, uint32_t w, uint32_t c) {
    const uint32_t bit = w * c, limb = bit >> 5, sh = bit & 31u;
    uint64_t v = canon.v[limb];
    if (limb + 1 < 8) v |= (uint64_t)canon.v[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

struct MsmDigitsBody {
    struct Args { const uint4* scalars; uint32_t* keys; uint32_t* vals; uint64_t n; uint32_t c, n_windows; };
    PG_HD static void run(const Args& a, uint64_t i) {
        const Fr s = fr_from_mont(aos_load(a.scalars, i));
        for (uint32_t w = 0; w < a.n_windows; w++) {
            a.keys[(uint64_t)w * a.n + i] = (w << a.c) | msm_digit(s, w, a.c);
            a.vals[(uint64_t)w * a.n + i] = (uint32_t)i;
        }
    }
};

// first position in sorted keys[0..count) whose key is >= key
PG_HD uint64_t msm_lower_bound(const uint32_t* keys, uint64_t count, uint32_t key) {
    uint64_t lo = 0, hi = count;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// run of bucket b in the sorted pairs: start and length (digit 0 buckets get length 0: they contribute nothing).  The length is
// also written as a sort key (descending) so that the bucket kernel can hand the lanes of a warp buckets of equal length.
struct MsmBoundsBody {
    struct Args { const uint32_t* keys; unsigned long long* start; uint32_t* size_key; uint32_t* ids; uint64_t count; uint64_t n; /* buckets */ uint32_t c; };
    PG_HD static void run(const Args& a, uint64_t b) {
        uint64_t lo = 0, len = 0;
        if (b & ((1ull << a.c) - 1ull)) {
            lo = msm_lower_bound(a.keys, a.count, (uint32_t)b);
            len = msm_lower_bound(a.keys, a.count, (uint32_t)b + 1u) - lo;
        }
        a.start[b] = lo;
        a.size_key[b] = 0xffffffffu - (uint32_t)(len > 0xffffffffull ? 0xffffffffull : len);
        a.ids[b] = (uint32_t)b;
    }
};

// Bucket sums in three levels so that no thread ever walks a long run, whatever the scalars look like (wire VALUES put millions of
// points into the "digit 1" bucket; uniform scalars give every bucket a few dozen):
//   parts  : a bucket's run is cut into ceil(len / MSM_PART) parts of equal length; parts are numbered bucket after bucket in
//            order of decreasing bucket length (the order of the small sort), so that neighbouring parts are equally long;
//   groups : a bucket's parts are cut into groups of at most MSM_GROUP parts;
//   finish : one thread per bucket adds its groups.
// Offsets come from two exclusive scans; the totals are bounded on the host (count / MSM_PART + buckets), surplus threads leave.
constexpr uint32_t MSM_PART = 512, MSM_GROUP = 64;
struct MsmPartsBody {           // per bucket, in sorted order: number of parts and of groups
    struct Args { const uint32_t* size_key; uint32_t* parts; uint32_t* groups; uint64_t n; };
    PG_HD static void run(const Args& a, uint64_t t) {
        const uint32_t len = 0xffffffffu - a.size_key[t];
        const uint32_t p = (len + MSM_PART - 1) / MSM_PART;
        a.parts[t] = p; a.groups[t] = (p + MSM_GROUP - 1) / MSM_GROUP;
    }
};
struct MsmExpandBody {          // owner (sorted bucket position) of every part and of every group
    struct Args { const uint32_t* parts; const uint32_t* groups; const uint32_t* off1; const uint32_t* off2; uint32_t* part_owner; uint32_t* group_owner; uint64_t n; };
    PG_HD static void run(const Args& a, uint64_t t) {
        for (uint32_t j = 0; j < a.parts[t]; j++) a.part_owner[a.off1[t] + j] = (uint32_t)t;
        for (uint32_t j = 0; j < a.groups[t]; j++) a.group_owner[a.off2[t] + j] = (uint32_t)t;
    }
};
struct MsmBucketBody {          // level 1: mixed additions of the points of one part into an XYZZ accumulator
    struct Args { const uint32_t* ids; const uint32_t* size_key; const unsigned long long* start; const uint32_t* vals; const uint4* points;
                  const uint32_t* parts; const uint32_t* off1; const uint32_t* part_owner; uint4* p1; uint64_t n; /* bound on parts */ uint64_t n_buckets; };
    PG_HD static void run(const Args& a, uint64_t p) {
        const uint64_t total = (uint64_t)a.off1[a.n_buckets - 1] + a.parts[a.n_buckets - 1];
        if (p >= total) return;
        const uint32_t t = a.part_owner[p], b = a.ids[t];
        const uint64_t len = 0xffffffffu - a.size_key[t], j = p - a.off1[t], np = a.parts[t];
        const uint64_t each = len / np, extra = len % np;                    // equal parts: the first `extra` ones hold one entry more
        const uint64_t lo = a.start[b] + j * each + (j < extra ? j : extra), hi = lo + each + (j < extra ? 1 : 0);
        G1X acc = g1x_inf();
        for (uint64_t k = lo; k < hi; k++) acc = g1x_madd(acc, g1_affine_load(a.points, a.vals[k]));
        g1x_store(a.p1, p, acc);
    }
};
struct MsmGroupSumBody {        // level 2: up to MSM_GROUP consecutive parts of one bucket
    struct Args { const uint32_t* parts; const uint32_t* groups; const uint32_t* off1; const uint32_t* off2; const uint32_t* group_owner; const uint4* p1; uint4* p2;
                  uint64_t n; /* bound on groups */ uint64_t n_buckets; };
    PG_HD static void run(const Args& a, uint64_t g) {
        const uint64_t total = (uint64_t)a.off2[a.n_buckets - 1] + a.groups[a.n_buckets - 1];
        if (g >= total) return;
        const uint32_t t = a.group_owner[g];
        const uint32_t j = (uint32_t)(g - a.off2[t]), first = j * MSM_GROUP;
        const uint32_t cnt = a.parts[t] - first < MSM_GROUP ? a.parts[t] - first : MSM_GROUP;
        G1X acc = g1x_inf();
        for (uint32_t k = 0; k < cnt; k++) acc = g1x_add(acc, g1x_load(a.p1, (uint64_t)a.off1[t] + first + k));
        g1x_store(a.p2, g, acc);
    }
};
struct MsmBucketFinishBody {    // level 3: the groups of one bucket
    struct Args { const uint32_t* ids; const uint32_t* groups; const uint32_t* off2; const uint4* p2; uint4* buckets; uint64_t n; };
    PG_HD static void run(const Args& a, uint64_t t) {
        G1X acc = g1x_inf();
        for (uint32_t k = 0; k < a.groups[t]; k++) acc = g1x_add(acc, g1x_load(a.p2, (uint64_t)a.off2[t] + k));
        g1x_store(a.buckets, a.ids[t], acc);
    }
};

// chunk t of the flattened (window, digit) space: sum_{d in chunk} d * B_d for its window
struct MsmChunkBody {
    struct Args { const uint4* buckets; uint4* out; uint64_t n; /* chunks */ uint32_t c, chunk; };
    PG_HD static void run(const Args& a, uint64_t t) {
        const uint64_t first = t * a.chunk;                    // flattened index of the chunk's first bucket
        const uint32_t lo = (uint32_t)(first & ((1ull << a.c) - 1ull));
        G1X running = g1x_inf(), acc = g1x_inf();
        for (uint32_t k = a.chunk; k-- > 0;) {                 // acc = sum (d - lo) * B_d, running = sum B_d
            acc = g1x_add(acc, running);
            running = g1x_add(running, g1x_load(a.buckets, first + k));
        }
        if (lo) { const uint32_t kk[1] = {lo}; acc = g1x_add(acc, g1x_mul_limbs(running, kk, 1)); }
        g1x_store(a.out, t, acc);
    }
};

// out[t] = sum of in[t*group .. t*group + group) (clipped to per-window segments of seg_in entries -> seg_out entries)
struct MsmSumBody {
    struct Args { const uint4* in; uint4* out; uint64_t n; /* outputs */ uint32_t seg_in, seg_out, group; };
    PG_HD static void run(const Args& a, uint64_t t) {
        const uint64_t w = t / a.seg_out, j = t % a.seg_out;
        G1X acc = g1x_inf();
        for (uint32_t k = 0; k < a.group; k++) {
            const uint64_t idx = j * a.group + k;
            if (idx < a.seg_in) acc = g1x_add(acc, g1x_load(a.in, w * a.seg_in + idx));
        }
        g1x_store(a.out, t, acc);
    }
};

// result[t] = sum_w 2^(c*w) * window[t][w], as an affine point.  One thread per MSM: the Horner chain (c * n_windows ~ 255 dependent
// doublings) is serial by nature, so MSMs that are computed together (the four wire commitments) finish in ONE launch -- their
// chains run side by side in one warp instead of one after the other.
struct MsmFinalBody {
    struct Args { const uint4* windows; uint4* out; uint64_t n; /* MSMs */ uint32_t c, n_windows; };
    PG_HD static void run(const Args& a, uint64_t t) {
        G1X total = g1x_inf();
        for (uint32_t w = a.n_windows; w-- > 0;) {
            for (uint32_t k = 0; k < a.c; k++) total = g1x_dbl(total);
            total = g1x_add(total, g1x_load(a.windows, t * a.n_windows + w));
        }
        g1_affine_store(a.out, t, g1x_to_affine(total));
    }
};

// out[i] = scalar_i * base  (PublicParameters::setup: powers_of_g = slow_multiscalar_mul_single_base(powers_of_beta, g))
struct G1FixedBaseMulBody {
    struct Args { const uint4* scalars; uint4* out; uint64_t n; G1Affine base; };
    PG_HD static void run(const Args& a, uint64_t i) {
        const Fr k = fr_from_mont(aos_load(a.scalars, i));
        g1_affine_store(a.out, i, g1x_to_affine<true>(g1x_mul_limbs(g1x_from_affine(a.base), k.v, 8)));
    }
};

// util::powers_of(beta, n): out[i] = beta^i.  Thread t owns i in [64t, 64t + 64): beta^(64t) by square-and-multiply, then 63 products.
constexpr uint64_t FRPOW_CHUNK = 64;
struct FrPowersBody {
    struct Args { Fr beta; uint4* out; uint64_t n; uint64_t n_points; };
    PG_HD static void run(const Args& a, uint64_t t) {
        Fr acc = fr_one(), sq = a.beta;
#pragma unroll 1
        for (uint64_t e = t * FRPOW_CHUNK; e; e >>= 1) { if (e & 1) acc = fr_mul(acc, sq); sq = fr_mul(sq, sq); }
        const uint64_t lo = t * FRPOW_CHUNK, hi = lo + FRPOW_CHUNK < a.n_points ? lo + FRPOW_CHUNK : a.n_points;
#pragma unroll 1
        for (uint64_t i = lo; i < hi; i++) { aos_store(a.out, i, acc); acc = fr_mul(acc, a.beta); }
    }
};

// The same through a windowed table of multiples of the base (used from FB_MIN_POINTS scalars on, FB_MIN_POINTS_CACHED when the table exists): table[w * 255 + d - 1] =
// d * 2^(8w) * base for the 32 byte-windows of a scalar, so that one scalar multiplication is at most 32 mixed additions instead of
// 255 doublings + ~128 additions; the XYZZ results are normalised to affine with Montgomery's trick (one inversion per FB_CHUNK points).
// The accumulated scalar stays below q < the group order at every step, so an addition never meets its own operand (no doubling
// case); the complete formulas cover it anyway.
constexpr uint32_t FB_WINDOWS = 32, FB_ENTRIES = 255, FB_CHUNK = 32;
constexpr uint64_t FB_MIN_POINTS = 65536, FB_MIN_POINTS_CACHED = 256, FB_TILE = 1ull << 20;    // building the table: ~10 ms (248 dependent doublings); the plain kernel does 3.1 M/s
struct G1WindowTableBody {
    struct Args { uint4* table; uint64_t n; G1Affine base; };
    PG_HD static void run(const Args& a, uint64_t t) {
        const uint32_t w = (uint32_t)(t / FB_ENTRIES), d = (uint32_t)(t % FB_ENTRIES) + 1u;
        G1X p = g1x_from_affine(a.base);
#pragma unroll 1
        for (uint32_t j = 0; j < 8u * w; j++) p = g1x_dbl(p);                       // 2^(8w) * base
        g1_affine_store(a.table, t, g1x_to_affine<true>(g1x_mul_limbs(p, &d, 1)));
    }
};
struct G1FixedBaseWindowedBody {
    struct Args { const uint4* scalars; const uint4* table; uint4* xyzz; uint64_t n; };
    PG_HD static void run(const Args& a, uint64_t i) {
        const Fr k = fr_from_mont(aos_load(a.scalars, i));
        G1X acc = g1x_inf();
#pragma unroll 1
        for (uint32_t w = 0; w < FB_WINDOWS; w++) {
            const uint32_t d = (k.v[w >> 2] >> (8u * (w & 3u))) & 255u;
            if (d) acc = g1x_madd(acc, g1_affine_load(a.table, (uint64_t)w * FB_ENTRIES + d - 1u));
        }
        g1x_store(a.xyzz, i, acc);
    }
};
// n_points XYZZ points -> affine.  Thread t owns the points t, t + n, t + 2n, ... (n = number of threads: neighbouring lanes touch
// neighbouring points); prefix products of Z^5 = zz*zzz are parked in `prefix`, one inversion per thread.
struct G1BatchAffineBody {
    struct Args { const uint4* xyzz; uint4* prefix; uint4* out; uint64_t n; uint64_t n_points; };
    PG_HD static void run(const Args& a, uint64_t t) {
        Fp p = fp_one();
        uint64_t last = t;
#pragma unroll 1
        for (uint64_t e = t; e < a.n_points; e += a.n) {
            const G1X P = g1x_load(a.xyzz, e);
            if (!g1x_is_inf(P)) p = fp_mul(p, fp_mul(P.zz, P.zzz));
            fp_store(a.prefix + 3 * e, p);
            last = e;
        }
        Fp inv = fp_inv_fermat(p);                                                   // fixed-length chain: the lanes stay in step
#pragma unroll 1
        for (uint64_t e = last;; e -= a.n) {
            const G1X P = g1x_load(a.xyzz, e);
            if (g1x_is_inf(P)) g1_affine_store(a.out, e, g1_affine_inf());
            else {
                const Fp prev = e >= a.n + t ? fp_load(a.prefix + 3 * (e - a.n)) : fp_one();
                const Fp i5 = fp_mul(inv, prev);                                     // (zz*zzz)^-1 of this point
                inv = fp_mul(inv, fp_mul(P.zz, P.zzz));
                G1Affine q;
                q.x = fp_mul(P.x, fp_mul(i5, P.zzz));                                // X / ZZ
                q.y = fp_mul(P.y, fp_mul(i5, P.zz));                                 // Y / ZZZ
                g1_affine_store(a.out, e, q);
            }
            if (e < a.n + t) break;
        }
    }
};

// Lagrange-basis SRS scalars: L_i(beta) = (beta^n - 1) / n * w^i / (beta - w^i), i < n, for the domain generated by w
// (w^i from the NTT twiddle table: w^(i + n/2) = -w^i).  With powers [L_i(beta)] G a commitment can be computed from the
// EVALUATIONS of a polynomial over the domain -- sum_i f(w^i) [L_i(beta)] G is the same group element as
// sum_j coeff_j [beta^j] G -- and wire values are mostly bits and short accumulators: most of their windows are empty.
struct LagrangeScalarsBody {
    struct Args { const uint4* tw; uint4* out; uint64_t n; uint32_t log_n; Fr beta, c; };
    PG_HD static void run(const Args& a, uint64_t i) {
        Fr w = fr_one();
        if (a.log_n) { const uint64_t h = a.n >> 1; w = i < h ? aos_load(a.tw, i) : fr_neg(aos_load(a.tw, i - h)); }
        const Fr d = fr_sub(a.beta, w);
        aos_store(a.out, i, fr_is_zero(d) ? fr_zero() : fr_mul(fr_mul(a.c, w), fr_inv_fermat(d)));   // beta on the domain is rejected by the caller
    }
};

// group-law self-test entry (pg_g1_op): 0 = a + b (affine in, affine out), 1 = on-curve flag of a
struct G1OpBody {
    struct Args { const uint4* a; const uint4* b; uint4* out; uint64_t n; int op; };
    PG_HD static void run(const Args& x, uint64_t i) {
        const G1Affine p = g1_affine_load(x.a, i);
        if (x.op == 0) { g1_affine_store(x.out, i, g1x_to_affine(g1x_madd(g1x_from_affine(p), g1_affine_load(x.b, i)))); return; }
        G1Affine r = g1_affine_inf();
        const Fp four = fp_dbl(fp_dbl(fp_one()));
        const bool ok = g1_affine_is_inf(p) || fp_eq(fp_sqr(p.y), fp_add(fp_mul(fp_sqr(p.x), p.x), four));
        r.x.v[0] = ok ? 1u : 0u;
        g1_affine_store(x.out, i, r);
    }
};

}  // namespace pg
