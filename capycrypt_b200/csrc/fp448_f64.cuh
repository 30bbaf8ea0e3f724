// fp448_f64.cuh -- GF(2^448 - 2^224 - 1) multiplication on the FP64 + integer-multiply pipes (B200: DFMA and IMAD both
// run at 64 lanes/clk/SM, IMAD.WIDE at half that and it blocks the ALU pipe while it issues).
//
// Same representation as fp448.cuh (16 x 28-bit limbs, Karatsuba over phi = 2^224, columns k and 8 + k finished
// together).  A limb product is NOT formed as a 64-bit integer.  For every output column two things are accumulated:
//   chain   a double that starts at C0 = 1.5 * 2^80 and receives every product with fma.rz: in the binade [2^80, 2^81) one
//           ulp is 2^28, so after each step the chain is the exact running sum rounded DOWN to a multiple of 2^28
//           (error of one step in [0, 2^28), for products of either sign)
//   low     the same products multiplied as 32-bit integers (IMAD, low word only): the exact sum mod 2^32
// With at most 16 products per column the accumulated rounding error E is < 2^32, so E = (low - chain) mod 2^32 and the
// exact column value is chain + E: 1 DFMA + 1 IMAD per limb product instead of one IMAD.WIDE (two issue slots that
// also stall the ALU pipe).
#pragma once
#include "fp448.cuh"

#if !defined(__CUDA_ARCH__)
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace capy {

#if defined(__CUDA_ARCH__)
CAPY_HD double f64_fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
CAPY_HD double f64_from_u32(uint32_t x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }
CAPY_HD uint32_t f64_lo(double d) { return (uint32_t)__double2loint(d); }
CAPY_HD uint32_t f64_hi(double d) { return (uint32_t)__double2hiint(d); }
#else
// host build (unit tests only): the caller runs under fesetround(FE_TOWARDZERO); compile with -frounding-math
CAPY_HD double f64_fma_rz(double a, double b, double c) { return std::fma(a, b, c); }
CAPY_HD double f64_from_u32(uint32_t x) { return (double)x; }
CAPY_HD uint32_t f64_lo(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (uint32_t)u; }
CAPY_HD uint32_t f64_hi(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (uint32_t)(u >> 32); }
#endif

// C0 = 1.5 * 2^80: bits 0x44F8000000000000
constexpr double F64_C0 = 1813388729421943762059264.0;
constexpr uint32_t F64_C0_HI = 0x44F80000u;

struct F64Chain {
  double d;     // C0 + (running sum rounded down to multiples of 2^28)
  uint32_t lo;  // running sum mod 2^32
};

CAPY_HD void ch_init(F64Chain& c) {
  c.d = F64_C0;
  c.lo = 0;
}
CAPY_HD void ch_mac(F64Chain& c, double da, double db, uint32_t a, uint32_t b) {
  c.d = f64_fma_rz(da, db, c.d);
  c.lo += a * b;
}
CAPY_HD void ch_msub(F64Chain& c, double da, double db, uint32_t a, uint32_t b) {
  c.d = f64_fma_rz(-da, db, c.d);
  c.lo -= a * b;
}

// Finish one column: exact value = 2^28 (h - hbias) + E with E in [0, 2^32), where h = chain mantissa relative to C0 and
// the accumulated rounding error lies in (-hbias 2^28, 2^32 - hbias 2^28).  Adds the carry, emits the 28-bit limb, returns
// the carry out (signed 64-bit; the running totals are non-negative).
CAPY_HD uint32_t ch_finish(const F64Chain& c, int hbias, int64_t& carry) {
  const uint32_t blo = f64_lo(c.d), bhi = f64_hi(c.d) - F64_C0_HI;  // h = (bhi : blo) as a signed 64-bit integer
  const int64_t h = (int64_t)(((uint64_t)bhi << 32) | blo) - hbias;
  const uint32_t E = c.lo - (blo << 28) + ((uint32_t)hbias << 28);  // (low - 2^28 h + bias) mod 2^32
  const int64_t w = carry + (int64_t)(uint64_t)E;
  carry = h + (w >> 28);
  return (uint32_t)w & M28;
}

// r = a * b.  Same contract as fe_mul_inl: alpha_a * alpha_b <= 6 (each alpha < 8), r may alias a or b.
CAPY_HD void fe_mul_f64(Fe& r, const Fe& a, const Fe& b) {
  uint32_t a0[8], a1[8], b0[8], b1[8], s[8], t[8];
  double A0[8], A1[8], B0[8], B1[8], S[8], T[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a0[i] = a.v[i];
    a1[i] = a.v[i + 8];
    b0[i] = b.v[i];
    b1[i] = b.v[i + 8];
    s[i] = a0[i] + a1[i];
    t[i] = b0[i] + b1[i];
    A0[i] = f64_from_u32(a0[i]);
    A1[i] = f64_from_u32(a1[i]);
    B0[i] = f64_from_u32(b0[i]);
    B1[i] = f64_from_u32(b1[i]);
    S[i] = A0[i] + A1[i];  // exact (< 2^33)
    T[i] = B0[i] + B1[i];
  }
  // with U = a0*b0, W = a1*b1, Y = s*t as 15-column products (lo = columns 0..7, hi = 8..14):
  //   r[k]     = Ulo[k] + Wlo[k] + Yhi[k] - Uhi[k]
  //   r[8 + k] = Whi[k] + Yhi[k] + Ylo[k] - Ulo[k]
  uint32_t v[16];
  int64_t clo = 0, chi = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    F64Chain z, u, lo, hi;
    ch_init(z);
    ch_init(u);
#pragma unroll
    for (int i = k + 1; i < 8; i++) ch_mac(z, S[i], T[k + 8 - i], s[i], t[k + 8 - i]);  // Yhi[k]
#pragma unroll
    for (int i = 0; i <= k; i++) ch_mac(u, A0[i], B0[k - i], a0[i], b0[k - i]);  // Ulo[k]
    const double tu = u.d - F64_C0;  // exact: a multiple of 2^28 far below 2^79
    lo.d = z.d + tu;
    hi.d = z.d - tu;
    lo.lo = z.lo + u.lo;
    hi.lo = z.lo - u.lo;
#pragma unroll
    for (int i = 0; i <= k; i++) ch_mac(lo, A1[i], B1[k - i], a1[i], b1[k - i]);  // Wlo[k]
#pragma unroll
    for (int i = k + 1; i < 8; i++) ch_msub(lo, A0[i], B0[k + 8 - i], a0[i], b0[k + 8 - i]);  // -Uhi[k]
#pragma unroll
    for (int i = k + 1; i < 8; i++) ch_mac(hi, A1[i], B1[k + 8 - i], a1[i], b1[k + 8 - i]);  // Whi[k]
#pragma unroll
    for (int i = 0; i <= k; i++) ch_mac(hi, S[i], T[k - i], s[i], t[k - i]);  // Ylo[k]
    // rounding errors: lo collected 16 steps, all >= 0; hi collected 15 - k steps >= 0 and subtracted the k + 1 of u
    v[k] = ch_finish(lo, 0, clo);
    v[8 + k] = ch_finish(hi, k + 1, chi);
  }
  fe_carry_tail(r, v, (uint64_t)clo, (uint64_t)chi);
}

// r = a^2.  Requires alpha_a^2 <= 6.
CAPY_HD void fe_sqr_f64(Fe& r, const Fe& a) {
  uint32_t a0[8], a1[8], s[8];
  double A0[8], A1[8], S[8], D0[8], D1[8], DS[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a0[i] = a.v[i];
    a1[i] = a.v[i + 8];
    s[i] = a0[i] + a1[i];
    A0[i] = f64_from_u32(a0[i]);
    A1[i] = f64_from_u32(a1[i]);
    S[i] = A0[i] + A1[i];
    D0[i] = A0[i] + A0[i];
    D1[i] = A1[i] + A1[i];
    DS[i] = S[i] + S[i];
  }
  // column m of x^2 = sum_{i<j, i+j=m} (2 x_i) x_j + [m even] x_{m/2}^2 ; the integer side doubles by a shift of the sum
  uint32_t v[16];
  int64_t clo = 0, chi = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    F64Chain z, u, lo, hi;
    ch_init(z);
    ch_init(u);
#pragma unroll
    for (int i = k + 1; i < 8; i++) {  // Yhi[k]: column k + 8 of s^2
      const int j = k + 8 - i;
      if (i < j) ch_mac(z, DS[i], S[j], 2u * s[i], s[j]);
      if (i == j) ch_mac(z, S[i], S[i], s[i], s[i]);
    }
#pragma unroll
    for (int i = 0; i <= k; i++) {  // Ulo[k]: column k of a0^2
      const int j = k - i;
      if (i < j) ch_mac(u, D0[i], A0[j], 2u * a0[i], a0[j]);
      if (i == j) ch_mac(u, A0[i], A0[i], a0[i], a0[i]);
    }
    const double tu = u.d - F64_C0;
    lo.d = z.d + tu;
    hi.d = z.d - tu;
    lo.lo = z.lo + u.lo;
    hi.lo = z.lo - u.lo;
#pragma unroll
    for (int i = 0; i <= k; i++) {  // Wlo[k]
      const int j = k - i;
      if (i < j) ch_mac(lo, D1[i], A1[j], 2u * a1[i], a1[j]);
      if (i == j) ch_mac(lo, A1[i], A1[i], a1[i], a1[i]);
    }
#pragma unroll
    for (int i = k + 1; i < 8; i++) {  // -Uhi[k]
      const int j = k + 8 - i;
      if (i < j) ch_msub(lo, D0[i], A0[j], 2u * a0[i], a0[j]);
      if (i == j) ch_msub(lo, A0[i], A0[i], a0[i], a0[i]);
    }
#pragma unroll
    for (int i = k + 1; i < 8; i++) {  // Whi[k]
      const int j = k + 8 - i;
      if (i < j) ch_mac(hi, D1[i], A1[j], 2u * a1[i], a1[j]);
      if (i == j) ch_mac(hi, A1[i], A1[i], a1[i], a1[i]);
    }
#pragma unroll
    for (int i = 0; i <= k; i++) {  // Ylo[k]
      const int j = k - i;
      if (i < j) ch_mac(hi, DS[i], S[j], 2u * s[i], s[j]);
      if (i == j) ch_mac(hi, S[i], S[i], s[i], s[i]);
    }
    v[k] = ch_finish(lo, 0, clo);
    v[8 + k] = ch_finish(hi, k + 1, chi);
  }
  fe_carry_tail(r, v, (uint64_t)clo, (uint64_t)chi);
}

}  // namespace capy
