// g1.cuh -- BLS12-381 base field Fp and the group G1 (y^2 = x^3 + 4), as far as KZG commitments need them
// (second half: `commit_key.commit(&w_l_poly)` = msm_variable_base(powers_of_g, coeffs)).
//
// Fp: 12 x u32 little-endian limbs of a*2^384 mod p, always fully reduced -- byte-identical to dusk-bls12_381's `Fp([u64; 6])`.
// Points cross the C ABI as pg_g1_affine { x, y } (96 bytes, Montgomery limbs); the point at infinity is the all-zero pair
// (not on the curve since b = 4 != 0).  Sums are kept in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0 marks
// infinity): a mixed addition costs 8 multiplications + 2 squarings, no inversion until the final conversion.
// Every formula handles its exceptional cases (equal points, opposite points, infinity) -- bucket sums of real polynomials
// rarely hit them, the parity tests do on purpose.
#pragma once
#include "fr.cuh"

namespace pg {

struct Fp { uint32_t v[12]; };
struct G1Affine { Fp x, y; };
struct G1X { Fp x, y, zz, zzz; };

constexpr uint32_t FP_INV32 = 0xfffcfffdu;     // -p^-1 mod 2^32
PG_HD uint32_t fp_p(int i) {
    const uint32_t P[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                            0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return P[i];
}
PG_HD Fp fp_zero() { Fp r; for (int i = 0; i < 12; i++) r.v[i] = 0; return r; }
PG_HD Fp fp_one() {   // 2^384 mod p
    Fp r = {{0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u}};
    return r;
}
PG_HD bool fp_is_zero(const Fp& a) { uint32_t o = 0; for (int i = 0; i < 12; i++) o |= a.v[i]; return o == 0; }
PG_HD bool fp_eq(const Fp& a, const Fp& b) { uint32_t o = 0; for (int i = 0; i < 12; i++) o |= a.v[i] ^ b.v[i]; return o == 0; }

// r in [0, 2p) -> r mod p
PG_HD Fp fp_reduce_once(const Fp& t) {
    Fp d; uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { const uint64_t s = (uint64_t)t.v[i] - fp_p(i) - bw; d.v[i] = (uint32_t)s; bw = (s >> 32) & 1u; }
    return bw ? t : d;
}
PG_HD Fp fp_add_generic(const Fp& a, const Fp& b) {
    Fp s; uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { const uint64_t t = (uint64_t)a.v[i] + b.v[i] + c; s.v[i] = (uint32_t)t; c = t >> 32; }
    return fp_reduce_once(s);                       // p < 2^381: no carry out of limb 11
}
PG_HD Fp fp_sub_generic(const Fp& a, const Fp& b) {
    Fp d; uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { const uint64_t s = (uint64_t)a.v[i] - b.v[i] - bw; d.v[i] = (uint32_t)s; bw = (s >> 32) & 1u; }
    const uint32_t mask = 0u - (uint32_t)bw;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { const uint64_t t = (uint64_t)d.v[i] + (fp_p(i) & mask) + c; d.v[i] = (uint32_t)t; c = t >> 32; }
    return d;
}

// CIOS Montgomery multiplication, 12 limb steps (each: 12 products a*b_i, m = t0 * (-p^-1), 12 products m*p): the host version
// and the definition the device version below is tested against (through the oracle) on the GPU
PG_HD Fp fp_mul_generic(const Fp& a, const Fp& b) {
    uint32_t t[14];
#pragma unroll
    for (int i = 0; i < 14; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) { const uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c; t[j] = (uint32_t)s; c = s >> 32; }
        uint64_t s = (uint64_t)t[12] + c; t[12] = (uint32_t)s; t[13] = (uint32_t)(s >> 32);
        const uint32_t m = t[0] * FP_INV32;
        c = ((uint64_t)m * fp_p(0) + t[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 12; j++) { const uint64_t u = (uint64_t)m * fp_p(j) + t[j] + c; t[j - 1] = (uint32_t)u; c = u >> 32; }
        s = (uint64_t)t[12] + c; t[11] = (uint32_t)s; t[12] = t[13] + (uint32_t)(s >> 32);
    }
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = t[i];
    return fp_reduce_once(r);                       // a, b < p => t < 2p < 2^382: t[12] == 0
}

#if defined(__CUDACC__)
__device__ __constant__ uint32_t c_p[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                            0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
#endif
#if defined(__CUDA_ARCH__)
// ---- device multiplier: the even/odd carry-chain scheme of fr.cuh (fr_mul_eo) for 12 limbs ---------------------------------
// Partial products are split into an even and an odd accumulator so that every row of 6 products is ONE
// mad.lo.cc / madc.hi.cc chain (-> IMAD.WIDE.U32.X); each primitive is one asm statement holding one complete chain.
// Bounds: the running value stays below p*(2^33 + 2) < 2^415, so the odd array (limbs 1..12) never carries out and the carry
// out of the even array (limbs 0..11) is absorbed by limb 12 = odd[11].  The modulus limbs come from constant memory as
// register operands (immediates would split the wide products, see fr.cuh).
PG_D void mul_row6(uint32_t* acc, const uint32_t* x, int off, uint32_t y) {
    asm("mul.lo.u32 %0, %12, %18;\n\tmul.hi.u32 %1, %12, %18;\n\t"
        "mul.lo.u32 %2, %13, %18;\n\tmul.hi.u32 %3, %13, %18;\n\t"
        "mul.lo.u32 %4, %14, %18;\n\tmul.hi.u32 %5, %14, %18;\n\t"
        "mul.lo.u32 %6, %15, %18;\n\tmul.hi.u32 %7, %15, %18;\n\t"
        "mul.lo.u32 %8, %16, %18;\n\tmul.hi.u32 %9, %16, %18;\n\t"
        "mul.lo.u32 %10, %17, %18;\n\tmul.hi.u32 %11, %17, %18;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]), "=&r"(acc[6]), "=&r"(acc[7]), "=&r"(acc[8]), "=&r"(acc[9]), "=&r"(acc[10]), "=&r"(acc[11])
        : "r"(x[off + 0]), "r"(x[off + 2]), "r"(x[off + 4]), "r"(x[off + 6]), "r"(x[off + 8]), "r"(x[off + 10]), "r"(y));
}

PG_D void mad_row6(uint32_t* acc, uint32_t& top, const uint32_t* x, int off, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %13, %19, %0;\n\tmadc.hi.cc.u32 %1, %13, %19, %1;\n\t"
        "madc.lo.cc.u32 %2, %14, %19, %2;\n\tmadc.hi.cc.u32 %3, %14, %19, %3;\n\t"
        "madc.lo.cc.u32 %4, %15, %19, %4;\n\tmadc.hi.cc.u32 %5, %15, %19, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %19, %6;\n\tmadc.hi.cc.u32 %7, %16, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %17, %19, %8;\n\tmadc.hi.cc.u32 %9, %17, %19, %9;\n\t"
        "madc.lo.cc.u32 %10, %18, %19, %10;\n\tmadc.hi.cc.u32 %11, %18, %19, %11;\n\t"
        "addc.u32 %12, %12, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(top)
        : "r"(x[off + 0]), "r"(x[off + 2]), "r"(x[off + 4]), "r"(x[off + 6]), "r"(x[off + 8]), "r"(x[off + 10]), "r"(y));
}

PG_D void mad_row6_nc(uint32_t* acc, const uint32_t* x, int off, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %12, %18, %0;\n\tmadc.hi.cc.u32 %1, %12, %18, %1;\n\t"
        "madc.lo.cc.u32 %2, %13, %18, %2;\n\tmadc.hi.cc.u32 %3, %13, %18, %3;\n\t"
        "madc.lo.cc.u32 %4, %14, %18, %4;\n\tmadc.hi.cc.u32 %5, %14, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %15, %18, %6;\n\tmadc.hi.cc.u32 %7, %15, %18, %7;\n\t"
        "madc.lo.cc.u32 %8, %16, %18, %8;\n\tmadc.hi.cc.u32 %9, %16, %18, %9;\n\t"
        "madc.lo.cc.u32 %10, %17, %18, %10;\n\tmadc.hi.u32 %11, %17, %18, %11;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
        : "r"(x[off + 0]), "r"(x[off + 2]), "r"(x[off + 4]), "r"(x[off + 6]), "r"(x[off + 8]), "r"(x[off + 10]), "r"(y));
}

PG_D void mad_row6_shift(uint32_t* o, uint32_t& e0, const uint32_t* x, int off, uint32_t y) {
    asm("add.cc.u32 %12, %12, %1;\n\t"
        "madc.lo.cc.u32 %0, %13, %19, %2;\n\tmadc.hi.cc.u32 %1, %13, %19, %3;\n\t"
        "madc.lo.cc.u32 %2, %14, %19, %4;\n\tmadc.hi.cc.u32 %3, %14, %19, %5;\n\t"
        "madc.lo.cc.u32 %4, %15, %19, %6;\n\tmadc.hi.cc.u32 %5, %15, %19, %7;\n\t"
        "madc.lo.cc.u32 %6, %16, %19, %8;\n\tmadc.hi.cc.u32 %7, %16, %19, %9;\n\t"
        "madc.lo.cc.u32 %8, %17, %19, %10;\n\tmadc.hi.cc.u32 %9, %17, %19, %11;\n\t"
        "madc.lo.cc.u32 %10, %18, %19, 0;\n\tmadc.hi.u32 %11, %18, %19, 0;"
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(e0)
        : "r"(x[off + 0]), "r"(x[off + 2]), "r"(x[off + 4]), "r"(x[off + 6]), "r"(x[off + 8]), "r"(x[off + 10]), "r"(y));
}

PG_D void merge_even_odd12(uint32_t* r, const uint32_t* e, const uint32_t* o) {
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, %33;\n\t"
        "addc.cc.u32 %10, %22, %34;\n\t"
        "addc.u32 %11, %23, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11])
        : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]));
}

PG_D void fp_red_rows(uint32_t* X, uint32_t* Y, const uint32_t* p) {
    uint32_t m;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(X[0]), "r"(FP_INV32));
    mad_row6_nc(Y, p, 1, m);
    mad_row6(X, Y[11], p, 0, m);
}
PG_D void fp_step_first(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi, const uint32_t* p) {
    mul_row6(Y, a, 1, bi);
    mul_row6(X, a, 0, bi);
    fp_red_rows(X, Y, p);
}
PG_D void fp_step(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi, const uint32_t* p) {
    mad_row6_shift(Y, X[0], a, 1, bi);          // X[0] += Y[1]; Y = (Y >> 64) + a_odd*bi
    mad_row6(X, Y[11], a, 0, bi);
    fp_red_rows(X, Y, p);
}
PG_D Fp fp_sub_p_if_ge(const Fp& t, const uint32_t* p) {      // t in [0, 2p) -> t mod p
    Fp d; uint32_t bw;
    asm("sub.cc.u32 %0, %13, %25;\n\tsubc.cc.u32 %1, %14, %26;\n\tsubc.cc.u32 %2, %15, %27;\n\tsubc.cc.u32 %3, %16, %28;\n\t"
        "subc.cc.u32 %4, %17, %29;\n\tsubc.cc.u32 %5, %18, %30;\n\tsubc.cc.u32 %6, %19, %31;\n\tsubc.cc.u32 %7, %20, %32;\n\t"
        "subc.cc.u32 %8, %21, %33;\n\tsubc.cc.u32 %9, %22, %34;\n\tsubc.cc.u32 %10, %23, %35;\n\tsubc.cc.u32 %11, %24, %36;\n\t"
        "subc.u32 %12, 0, 0;"
        : "=&r"(d.v[0]), "=&r"(d.v[1]), "=&r"(d.v[2]), "=&r"(d.v[3]), "=&r"(d.v[4]), "=&r"(d.v[5]), "=&r"(d.v[6]), "=&r"(d.v[7]),
          "=&r"(d.v[8]), "=&r"(d.v[9]), "=&r"(d.v[10]), "=&r"(d.v[11]), "=&r"(bw)
        : "r"(t.v[0]), "r"(t.v[1]), "r"(t.v[2]), "r"(t.v[3]), "r"(t.v[4]), "r"(t.v[5]), "r"(t.v[6]), "r"(t.v[7]), "r"(t.v[8]), "r"(t.v[9]), "r"(t.v[10]), "r"(t.v[11]),
          "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(p[8]), "r"(p[9]), "r"(p[10]), "r"(p[11]));
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = bw ? t.v[i] : d.v[i];
    return r;
}

// device add / sub: one carry chain each (the 64-bit emulation above costs three instructions per limb)
PG_D Fp fp_add(const Fp& a, const Fp& b) {
    Fp s;
    asm("add.cc.u32 %0, %12, %24;\n\taddc.cc.u32 %1, %13, %25;\n\taddc.cc.u32 %2, %14, %26;\n\taddc.cc.u32 %3, %15, %27;\n\taddc.cc.u32 %4, %16, %28;\n\taddc.cc.u32 %5, %17, %29;\n\taddc.cc.u32 %6, %18, %30;\n\taddc.cc.u32 %7, %19, %31;\n\taddc.cc.u32 %8, %20, %32;\n\taddc.cc.u32 %9, %21, %33;\n\taddc.cc.u32 %10, %22, %34;\n\taddc.u32 %11, %23, %35;"
        : "=&r"(s.v[0]), "=&r"(s.v[1]), "=&r"(s.v[2]), "=&r"(s.v[3]), "=&r"(s.v[4]), "=&r"(s.v[5]), "=&r"(s.v[6]), "=&r"(s.v[7]), "=&r"(s.v[8]), "=&r"(s.v[9]), "=&r"(s.v[10]), "=&r"(s.v[11])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
    uint32_t p[12];
#pragma unroll
    for (int i = 0; i < 12; i++) p[i] = c_p[i];
    return fp_sub_p_if_ge(s, p);
}
PG_D Fp fp_sub(const Fp& a, const Fp& b) {
    Fp d; uint32_t bw;
    asm("sub.cc.u32 %0, %13, %25;\n\tsubc.cc.u32 %1, %14, %26;\n\tsubc.cc.u32 %2, %15, %27;\n\tsubc.cc.u32 %3, %16, %28;\n\tsubc.cc.u32 %4, %17, %29;\n\tsubc.cc.u32 %5, %18, %30;\n\tsubc.cc.u32 %6, %19, %31;\n\tsubc.cc.u32 %7, %20, %32;\n\tsubc.cc.u32 %8, %21, %33;\n\tsubc.cc.u32 %9, %22, %34;\n\tsubc.cc.u32 %10, %23, %35;\n\tsubc.cc.u32 %11, %24, %36;\n\tsubc.u32 %12, 0, 0;"
        : "=&r"(d.v[0]), "=&r"(d.v[1]), "=&r"(d.v[2]), "=&r"(d.v[3]), "=&r"(d.v[4]), "=&r"(d.v[5]), "=&r"(d.v[6]), "=&r"(d.v[7]), "=&r"(d.v[8]), "=&r"(d.v[9]), "=&r"(d.v[10]), "=&r"(d.v[11]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
    Fp r;
    asm("add.cc.u32 %0, %12, %24;\n\taddc.cc.u32 %1, %13, %25;\n\taddc.cc.u32 %2, %14, %26;\n\taddc.cc.u32 %3, %15, %27;\n\taddc.cc.u32 %4, %16, %28;\n\taddc.cc.u32 %5, %17, %29;\n\taddc.cc.u32 %6, %18, %30;\n\taddc.cc.u32 %7, %19, %31;\n\taddc.cc.u32 %8, %20, %32;\n\taddc.cc.u32 %9, %21, %33;\n\taddc.cc.u32 %10, %22, %34;\n\taddc.u32 %11, %23, %35;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(r.v[8]), "=&r"(r.v[9]), "=&r"(r.v[10]), "=&r"(r.v[11])
        : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]), "r"(d.v[7]), "r"(d.v[8]), "r"(d.v[9]), "r"(d.v[10]), "r"(d.v[11]), "r"(c_p[0] & bw), "r"(c_p[1] & bw), "r"(c_p[2] & bw), "r"(c_p[3] & bw), "r"(c_p[4] & bw), "r"(c_p[5] & bw), "r"(c_p[6] & bw), "r"(c_p[7] & bw), "r"(c_p[8] & bw), "r"(c_p[9] & bw), "r"(c_p[10] & bw), "r"(c_p[11] & bw));
    return r;
}
PG_D Fp fp_mul(const Fp& a, const Fp& b) {
    uint32_t p[12];
#pragma unroll
    for (int i = 0; i < 12; i++) p[i] = c_p[i];
    uint32_t even[12], odd[12];
    fp_step_first(even, odd, a.v, b.v[0], p);
#pragma unroll
    for (int i = 1; i < 12; i += 2) {
        fp_step(odd, even, a.v, b.v[i], p);
        if (i + 1 < 12) fp_step(even, odd, a.v, b.v[i + 1], p);
    }
    // 12 steps: the last one ran with X = odd, Y = even, so the value is odd[1..11] (limbs 0..10) + even (limbs 0..11)
    Fp r;
    merge_even_odd12(r.v, even, odd);
    return fp_sub_p_if_ge(r, p);
}
#else
PG_HD Fp fp_mul(const Fp& a, const Fp& b) { return fp_mul_generic(a, b); }
PG_HD Fp fp_add(const Fp& a, const Fp& b) { return fp_add_generic(a, b); }
PG_HD Fp fp_sub(const Fp& a, const Fp& b) { return fp_sub_generic(a, b); }
#endif
PG_HD Fp fp_neg(const Fp& a) { return fp_sub(fp_zero(), a); }
PG_HD Fp fp_dbl(const Fp& a) { return fp_add(a, a); }
PG_HD Fp fp_sqr(const Fp& a) { return fp_mul(a, a); }
PG_HD Fp fp_to_mont(const Fp& raw) {
    const Fp r2 = {{0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u}};
    return fp_mul(raw, r2);
}
// a^(p-2); a != 0.  Kept as the independent check of fp_inv (tests/emu): 381 squarings + ~190 multiplications, each one
// dependent on the last.
PG_HD Fp fp_inv_fermat(const Fp& a) {
    Fp res = fp_one();
#pragma unroll 1
    for (int i = 11; i >= 0; i--) {
        const uint32_t e = i == 0 ? fp_p(0) - 2u : fp_p(i);
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            res = fp_sqr(res);
            if ((e >> b) & 1u) res = fp_mul(res, a);
        }
    }
    return res;
}
// The same inverse (it is unique) by the binary extended Euclidean algorithm, like fr_inv_binary (fr.cuh): ~550 short
// shift / add / subtract steps instead of ~570 dependent 12-limb multiplications -- the single-thread tail of every MSM ends
// with one of these (g1x_to_affine).  Invariant s*a == u, t*a == v (mod p) on the stored integers a~ = a R; at u == v == 1
// t = a~^-1 and mont(t, R^3) = a^-1 R.  a != 0; variable time.
PG_HD uint32_t fp_sub_limbs12(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // r = a - b, returns the borrow
    uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { uint64_t t = (uint64_t)a[i] - b[i] - bw; r[i] = (uint32_t)t; bw = (t >> 32) & 1; }
    return (uint32_t)bw;
}
PG_HD void fp_shr1(uint32_t* u) {
#pragma unroll
    for (int i = 0; i < 11; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
    u[11] >>= 1;
}
PG_HD void fp_halve_mod_p(Fp& s) {           // s/2 mod p: (s + p)/2 when s is odd (s + p < 2^382)
    const uint32_t odd = 0u - (s.v[0] & 1u);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { c += (uint64_t)s.v[i] + (fp_p(i) & odd); s.v[i] = (uint32_t)c; c >>= 32; }
    fp_shr1(s.v);
}
PG_HD Fp fp_inv(const Fp& a) {
    const Fp r3 = {{0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au, 0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u}};   // 2^1152 mod p
    Fp u = a, v, s = fp_zero(), t = fp_zero(), d;
#pragma unroll
    for (int i = 0; i < 12; i++) v.v[i] = fp_p(i);
    s.v[0] = 1;
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        while (!(u.v[0] & 1u)) { fp_shr1(u.v); fp_halve_mod_p(s); }
#pragma unroll 1
        while (!(v.v[0] & 1u)) { fp_shr1(v.v); fp_halve_mod_p(t); }
        if (fp_sub_limbs12(d.v, u.v, v.v)) {                  // u < v
            fp_sub_limbs12(v.v, v.v, u.v);
            t = fp_sub(t, s);
        } else {
            u = d;
            if (fp_is_zero(u)) break;                         // u == v == 1
            s = fp_sub(s, t);
        }
    }
    return fp_mul(t, r3);
}

// ---- G1 ----------------------------------------------------------------------------------------------------------------
PG_HD G1Affine g1_generator() {
    G1Affine g = {{{0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u, 0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u}},
                  {{0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u, 0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu}}};
    return g;
}
PG_HD bool g1_affine_is_inf(const G1Affine& a) { return fp_is_zero(a.x) && fp_is_zero(a.y); }
PG_HD G1Affine g1_affine_inf() { G1Affine a; a.x = fp_zero(); a.y = fp_zero(); return a; }
PG_HD G1X g1x_inf() { G1X r; r.x = fp_one(); r.y = fp_one(); r.zz = fp_zero(); r.zzz = fp_zero(); return r; }
PG_HD bool g1x_is_inf(const G1X& a) { return fp_is_zero(a.zz); }
PG_HD G1X g1x_from_affine(const G1Affine& a) {
    if (g1_affine_is_inf(a)) return g1x_inf();
    G1X r; r.x = a.x; r.y = a.y; r.zz = fp_one(); r.zzz = fp_one(); return r;
}
// 2 * (affine point)
PG_HD G1X g1x_dbl_affine(const G1Affine& a) {
    if (g1_affine_is_inf(a) || fp_is_zero(a.y)) return g1x_inf();
    const Fp u = fp_dbl(a.y), v = fp_sqr(u), w = fp_mul(u, v), s = fp_mul(a.x, v);
    const Fp x2 = fp_sqr(a.x), m = fp_add(fp_dbl(x2), x2);
    G1X r;
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, a.y));
    r.zz = v; r.zzz = w;
    return r;
}
PG_HD G1X g1x_dbl(const G1X& p) {
    if (g1x_is_inf(p) || fp_is_zero(p.y)) return g1x_inf();
    const Fp u = fp_dbl(p.y), v = fp_sqr(u), w = fp_mul(u, v), s = fp_mul(p.x, v);
    const Fp x2 = fp_sqr(p.x), m = fp_add(fp_dbl(x2), x2);
    G1X r;
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = fp_mul(v, p.zz); r.zzz = fp_mul(w, p.zzz);
    return r;
}
// p + q, q affine
PG_HD G1X g1x_madd(const G1X& p, const G1Affine& q) {
    if (g1_affine_is_inf(q)) return p;
    if (g1x_is_inf(p)) return g1x_from_affine(q);
    const Fp u2 = fp_mul(q.x, p.zz), s2 = fp_mul(q.y, p.zzz);
    const Fp pp_ = fp_sub(u2, p.x), r_ = fp_sub(s2, p.y);
    if (fp_is_zero(pp_)) return fp_is_zero(r_) ? g1x_dbl_affine(q) : g1x_inf();
    const Fp pp = fp_sqr(pp_), ppp = fp_mul(pp_, pp), qq = fp_mul(p.x, pp);
    G1X r;
    r.x = fp_sub(fp_sub(fp_sqr(r_), ppp), fp_dbl(qq));
    r.y = fp_sub(fp_mul(r_, fp_sub(qq, r.x)), fp_mul(p.y, ppp));
    r.zz = fp_mul(p.zz, pp); r.zzz = fp_mul(p.zzz, ppp);
    return r;
}
PG_HD G1X g1x_add(const G1X& p, const G1X& q) {
    if (g1x_is_inf(q)) return p;
    if (g1x_is_inf(p)) return q;
    const Fp u1 = fp_mul(p.x, q.zz), u2 = fp_mul(q.x, p.zz), s1 = fp_mul(p.y, q.zzz), s2 = fp_mul(q.y, p.zzz);
    const Fp pp_ = fp_sub(u2, u1), r_ = fp_sub(s2, s1);
    if (fp_is_zero(pp_)) return fp_is_zero(r_) ? g1x_dbl(p) : g1x_inf();
    const Fp pp = fp_sqr(pp_), ppp = fp_mul(pp_, pp), qq = fp_mul(u1, pp);
    G1X r;
    r.x = fp_sub(fp_sub(fp_sqr(r_), ppp), fp_dbl(qq));
    r.y = fp_sub(fp_mul(r_, fp_sub(qq, r.x)), fp_mul(s1, ppp));
    r.zz = fp_mul(fp_mul(p.zz, q.zz), pp); r.zzz = fp_mul(fp_mul(p.zzz, q.zzz), ppp);
    return r;
}
// LOCKSTEP: every lane of the warp converts a point of its own -- the fixed-length Fermat chain keeps them in step, the
// data-dependent Euclid loop would not.
template <bool LOCKSTEP = false>
PG_HD G1Affine g1x_to_affine(const G1X& p) {
    if (g1x_is_inf(p)) return g1_affine_inf();
    const Fp zp = fp_mul(p.zz, p.zzz);
    const Fp i5 = LOCKSTEP ? fp_inv_fermat(zp) : fp_inv(zp);   // Z^-5
    G1Affine a;
    a.x = fp_mul(p.x, fp_mul(i5, p.zzz));                 // X / ZZ
    a.y = fp_mul(p.y, fp_mul(i5, p.zz));                  // Y / ZZZ
    return a;
}
// k * p for a canonical (non-Montgomery) 256-bit k given as 8 limbs, most significant bit first
PG_HD G1X g1x_mul_limbs(const G1X& p, const uint32_t* k, int n_limbs) {
    G1X acc = g1x_inf();
    bool started = false;
#pragma unroll 1
    for (int i = n_limbs - 1; i >= 0; i--) {
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            if (started) acc = g1x_dbl(acc);
            if ((k[i] >> b) & 1u) { acc = g1x_add(acc, p); started = true; }
        }
    }
    return acc;
}

// AoS access to caller memory / scratch (96-byte affine points, 192-byte XYZZ points), 16-byte units
PG_HD Fp fp_load(const uint4* p) {
    Fp r;
#pragma unroll
    for (int k = 0; k < 3; k++) { const uint4 q = p[k]; r.v[4 * k] = q.x; r.v[4 * k + 1] = q.y; r.v[4 * k + 2] = q.z; r.v[4 * k + 3] = q.w; }
    return r;
}
PG_HD void fp_store(uint4* p, const Fp& a) {
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = make_uint4(a.v[4 * k], a.v[4 * k + 1], a.v[4 * k + 2], a.v[4 * k + 3]);
}
PG_HD G1Affine g1_affine_load(const uint4* base, uint64_t i) { G1Affine a; a.x = fp_load(base + 6 * i); a.y = fp_load(base + 6 * i + 3); return a; }
PG_HD void g1_affine_store(uint4* base, uint64_t i, const G1Affine& a) { fp_store(base + 6 * i, a.x); fp_store(base + 6 * i + 3, a.y); }
PG_HD G1X g1x_load(const uint4* base, uint64_t i) {
    G1X r; r.x = fp_load(base + 12 * i); r.y = fp_load(base + 12 * i + 3); r.zz = fp_load(base + 12 * i + 6); r.zzz = fp_load(base + 12 * i + 9); return r;
}
PG_HD void g1x_store(uint4* base, uint64_t i, const G1X& a) {
    fp_store(base + 12 * i, a.x); fp_store(base + 12 * i + 3, a.y); fp_store(base + 12 * i + 6, a.zz); fp_store(base + 12 * i + 9, a.zzz);
}

}  // namespace pg
