// tests/emu/fr_emu.cpp -- host build of plonk_gadgets_b200/csrc/fr.cuh (TEST ONLY).
// The device multiplier's even/odd carry-chain algorithm is compiled here through its host emulation so that its limb
// logic (and the "dropped carry is zero" claims) can be checked against big-int arithmetic without a GPU.
#define PG_EMU_CHECKS 1
#include "../../plonk_gadgets_b200/csrc/fr.cuh"
#include "../../plonk_gadgets_b200/csrc/g1.cuh"
#include <cstdio>
#include <cstdlib>

static int g_violations = 0;
void pg_emu_carry_violation() { g_violations++; }

using pg::Fr;
extern "C" {
int emu_violations() { return g_violations; }
void emu_mul_eo(uint64_t n, const Fr* a, const Fr* b, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_mul_eo(a[i], b[i]); }
void emu_mul_cios(uint64_t n, const Fr* a, const Fr* b, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_mul_cios(a[i], b[i]); }
void emu_add(uint64_t n, const Fr* a, const Fr* b, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_add(a[i], b[i]); }
void emu_sub(uint64_t n, const Fr* a, const Fr* b, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_sub(a[i], b[i]); }
void emu_neg(uint64_t n, const Fr* a, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_neg(a[i]); }
void emu_inv(uint64_t n, const Fr* a, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_inv_fermat(a[i]); }
void emu_inv_binary(uint64_t n, const Fr* a, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_inv_binary(a[i]); }
void emu_fp_inv(uint64_t n, const pg::Fp* a, pg::Fp* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fp_inv(a[i]); }
void emu_fp_inv_fermat(uint64_t n, const pg::Fp* a, pg::Fp* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fp_inv_fermat(a[i]); }
void emu_from_mont(uint64_t n, const Fr* a, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_from_mont(a[i]); }
// r (9 limbs) = dot product of K=5 pairs through fr_dot_wide; ok[i] = limbs9_is_multiple_of_q(r + c)
void emu_dot5(uint64_t n, const Fr* a, const Fr* b, const Fr* c, uint32_t* r9, uint8_t* is_mult) {
    for (uint64_t i = 0; i < n; i++) {
        uint32_t r[9];
        pg::fr_dot_wide<5>(r, a + 5 * i, b + 5 * i);
        pg::add9_fr(r, c[i]);
        for (int k = 0; k < 9; k++) r9[9 * i + k] = r[k];
        is_mult[i] = pg::limbs9_is_multiple_of_q(r) ? 1 : 0;
    }
}
void emu_to_mont(uint64_t n, const Fr* a, Fr* r) { for (uint64_t i = 0; i < n; i++) r[i] = pg::fr_to_mont(a[i]); }
}
