// fp448.cuh -- GF(p), p = 2^448 - 2^224 - 1 (Ed448-Goldilocks), for sm_100a.
//
// Replaces, for batches, the field arithmetic the reference gets from the un-vendored crate
// tiny_ed448_goldilocks 0.1.8 (fiat-crypto p448_solinas_64; call sites ecc/keypair.rs:44,
// ecc/signable.rs:48-49,77-79, ecc/encryptable.rs:37-38,78).  Restated from the mathematics
// (SURVEY.md App. C.1); same values, different representation:
//
//   16 limbs x 28 bits in u32, unsaturated.  phi = 2^224 sits on limb 8, phi^2 = phi + 1, so a
//   product is one level of Karatsuba over phi: 3 x (8x8) = 192 multiply-accumulates.  Every
//   MAC is a single IMAD.WIDE (32x32 + 64 -> 64): products are < 2^60 and a column of them fits
//   64 bits, so there are no carry flags in the MAC phase.  Measured on B200 IMAD.WIDE is
//   half-rate (two issue slots, profiles/r01_peaks_int_pipes.json), which is why Karatsuba
//   (192 MACs + ~180 ALU ops) beats schoolbook (256 MACs).
//
// Bound discipline ("alpha" = max limb / 2^28):
//   tight  alpha <= 1 + 2^-20   every output of fe_mul / fe_sqr / fe_weak
//   fe_add of two tight -> alpha 2 ; fe_sub (a - b + 2p) of two tight -> alpha 3
//   fe_mul / fe_sqr accept alpha_a * alpha_b <= 6 (column sums stay below 2^64).
//   (also each alpha < 8: a0 + a1 must fit 32 bits and the negated a0 limbs a signed 32-bit operand)
// The functions are __host__ __device__ so the same code is unit-tested on the CPU
// (tests/host/ed448_host_check.cpp), which is test scaffolding, not a fallback.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CAPY_HD __host__ __device__ __forceinline__
#define CAPY_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CAPY_HD inline
#define CAPY_HD_NOINLINE
struct uint4 {  // host-side stand-in for the CUDA vector type (CPU unit tests only)
  uint32_t x, y, z, w;
};
#endif

namespace capy {

constexpr uint32_t M28 = (1u << 28) - 1u;

struct alignas(16) Fe {
  uint32_t v[16];
};

CAPY_HD void fe_zero(Fe& r) {
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = 0;
}
CAPY_HD void fe_one(Fe& r) {
  fe_zero(r);
  r.v[0] = 1;
}
CAPY_HD void fe_copy(Fe& r, const Fe& a) {
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = a.v[i];
}

// r = a + b, no carry (alpha adds)
CAPY_HD void fe_add(Fe& r, const Fe& a, const Fe& b) {
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = a.v[i] + b.v[i];
}

// r = a - b + 2p, no carry; requires b tight.  alpha_r = alpha_a + 2
CAPY_HD void fe_sub(Fe& r, const Fe& a, const Fe& b) {
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = a.v[i] + (i == 8 ? 2u * (M28 - 1u) : 2u * M28) - b.v[i];
}

// r = a - b + 4p, no carry; requires alpha_b <= 3.  alpha_r = alpha_a + 4
CAPY_HD void fe_sub4(Fe& r, const Fe& a, const Fe& b) {
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = a.v[i] + (i == 8 ? 4u * (M28 - 1u) : 4u * M28) - b.v[i];
}

// one parallel carry pass: any alpha < 16 -> tight
CAPY_HD void fe_weak(Fe& r) {
  uint32_t c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    c[i] = r.v[i] >> 28;
    r.v[i] &= M28;
  }
#pragma unroll
  for (int i = 1; i < 16; i++) r.v[i] += c[i - 1];
  r.v[0] += c[15];  // 2^448 = 2^224 + 1
  r.v[8] += c[15];
}

CAPY_HD void fe_neg(Fe& r, const Fe& a) {  // a tight -> tight
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = (i == 8 ? 2u * (M28 - 1u) : 2u * M28) - a.v[i];
  fe_weak(r);
}

// carry 16 64-bit columns to tight limbs: split every column in 28|28|8 bit pieces, add the
// pieces of neighbouring columns (no serial chain), then one parallel pass.
CAPY_HD void fe_carry_wide(Fe& r, const uint64_t (&R)[16]) {
  uint32_t p0[16], p1[16], p2[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t lo = (uint32_t)R[i], hi = (uint32_t)(R[i] >> 32);
    p0[i] = lo & M28;
    p1[i] = ((lo >> 28) | (hi << 4)) & M28;
    p2[i] = hi >> 24;
  }
  uint32_t v[16];
  v[0] = p0[0] + p1[15] + p2[14];
  v[1] = p0[1] + p1[0] + p2[15];
#pragma unroll
  for (int i = 2; i < 16; i++) v[i] = p0[i] + p1[i - 1] + p2[i - 2];
  v[8] += p1[15] + p2[14];
  v[9] += p2[15];
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = v[i];
  fe_weak(r);
}

// end of the two in-order carry chains of fe_mul / fe_sqr: v = limbs < 2^28, clo = carry out of limb 7 (weight
// 2^224, goes to limb 8), chi = carry out of limb 15 (weight 2^448 = 2^224 + 1, goes to limbs 0 and 8); both
// < 2^37.  One more step on limbs 0 and 8 leaves every limb below 2^28 + 2^10 (tight).
CAPY_HD void fe_carry_tail(Fe& r, uint32_t (&v)[16], uint64_t clo, uint64_t chi) {
  const uint64_t t0 = (uint64_t)v[0] + chi;
  const uint64_t t8 = (uint64_t)v[8] + clo + chi;
  v[0] = (uint32_t)t0 & M28;
  v[1] += (uint32_t)(t0 >> 28);
  v[8] = (uint32_t)t8 & M28;
  v[9] += (uint32_t)(t8 >> 28);
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = v[i];
}

// widening multiplies: unsigned for sums of limbs (may reach 2^32 - 1), signed only for the
// negated a0 limbs (|.| < 2^31).  Both are one IMAD.WIDE; accumulation is mod 2^64.
CAPY_HD uint64_t mulu(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
CAPY_HD uint64_t muls(int32_t na, uint32_t b) { return (uint64_t)((int64_t)na * (int64_t)(int32_t)b); }

// r = a * b.  Requires alpha_a * alpha_b <= 6 (and each alpha < 8).  r may alias a or b.
CAPY_HD void fe_mul_inl(Fe& r, const Fe& a, const Fe& b) {
  uint32_t a0[8], a1[8], b0[8], b1[8], s[8], t[8];
  int32_t na0[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a0[i] = a.v[i];
    a1[i] = a.v[i + 8];
    b0[i] = b.v[i];
    b1[i] = b.v[i + 8];
    s[i] = a0[i] + a1[i];
    t[i] = b0[i] + b1[i];
    na0[i] = -(int32_t)a0[i];
  }
  // with U = a0*b0, W = a1*b1, Y = s*t as 15-column products (lo = columns 0..7, hi = 8..14):
  //   r[k]     = Ulo[k] + Wlo[k] + Yhi[k] - Uhi[k]
  //   r[8 + k] = Whi[k] + Yhi[k] + Ylo[k] - Ulo[k]
  // Columns k and 8 + k are finished together and in order, so the carry of the two limb chains 0..7 and
  // 8..15 is resolved on the fly: one AND and one 64-bit shift per column, no separate carry pass.
  uint32_t v[16];
  uint64_t clo = 0, chi = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint64_t z = 0, u = 0;
#pragma unroll
    for (int i = k + 1; i < 8; i++) z += mulu(s[i], t[k + 8 - i]);  // Yhi[k]
#pragma unroll
    for (int i = 0; i <= k; i++) u += mulu(a0[i], b0[k - i]);  // Ulo[k]
    uint64_t lo = z + u + clo, hi = z - u + chi;  // the carries ride on the additions that are needed anyway
#pragma unroll
    for (int i = 0; i <= k; i++) lo += mulu(a1[i], b1[k - i]);  // Wlo[k]
#pragma unroll
    for (int i = k + 1; i < 8; i++) lo += muls(na0[i], b0[k + 8 - i]);  // -Uhi[k]
#pragma unroll
    for (int i = k + 1; i < 8; i++) hi += mulu(a1[i], b1[k + 8 - i]);  // Whi[k]
#pragma unroll
    for (int i = 0; i <= k; i++) hi += mulu(s[i], t[k - i]);  // Ylo[k]
    v[k] = (uint32_t)lo & M28;
    clo = lo >> 28;
    v[8 + k] = (uint32_t)hi & M28;
    chi = hi >> 28;
  }
  fe_carry_tail(r, v, clo, chi);
}

// r = a^2.  Requires alpha_a^2 <= 6.
CAPY_HD void fe_sqr_inl(Fe& r, const Fe& a) {
  uint32_t a0[8], a1[8], s[8], d0[8], d1[8], ds[8];
  int32_t nd0[8], na0[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a0[i] = a.v[i];
    a1[i] = a.v[i + 8];
    s[i] = a0[i] + a1[i];
    d0[i] = 2u * a0[i];
    d1[i] = 2u * a1[i];
    ds[i] = 2u * s[i];
    nd0[i] = -(int32_t)d0[i];
    na0[i] = -(int32_t)a0[i];
  }
  // column m of x^2 = sum_{i<j, i+j=m} (2 x_i) x_j + [m even] x_{m/2}^2
  uint32_t v[16];
  uint64_t clo = 0, chi = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint64_t z = 0, u = 0;
#pragma unroll
    for (int i = k + 1; i < 8; i++) {  // Yhi[k]: column k + 8 of s^2
      const int j = k + 8 - i;
      if (i < j) z += mulu(ds[i], s[j]);
      if (i == j) z += mulu(s[i], s[i]);
    }
#pragma unroll
    for (int i = 0; i <= k; i++) {  // Ulo[k]: column k of a0^2
      const int j = k - i;
      if (i < j) u += mulu(d0[i], a0[j]);
      if (i == j) u += mulu(a0[i], a0[i]);
    }
    uint64_t lo = z + u + clo, hi = z - u + chi;
#pragma unroll
    for (int i = 0; i <= k; i++) {  // Wlo[k]
      const int j = k - i;
      if (i < j) lo += mulu(d1[i], a1[j]);
      if (i == j) lo += mulu(a1[i], a1[i]);
    }
#pragma unroll
    for (int i = k + 1; i < 8; i++) {  // -Uhi[k]
      const int j = k + 8 - i;
      if (i < j) lo += muls(nd0[i], a0[j]);
      if (i == j) lo += muls(na0[i], a0[i]);
    }
#pragma unroll
    for (int i = k + 1; i < 8; i++) {  // Whi[k]
      const int j = k + 8 - i;
      if (i < j) hi += mulu(d1[i], a1[j]);
      if (i == j) hi += mulu(a1[i], a1[i]);
    }
#pragma unroll
    for (int i = 0; i <= k; i++) {  // Ylo[k]
      const int j = k - i;
      if (i < j) hi += mulu(ds[i], s[j]);
      if (i == j) hi += mulu(s[i], s[i]);
    }
    v[k] = (uint32_t)lo & M28;
    clo = lo >> 28;
    v[8 + k] = (uint32_t)hi & M28;
    chi = hi >> 28;
  }
  fe_carry_tail(r, v, clo, chi);
}

// out-of-line copies: one body instead of ~450 inlined instructions per call site.  Always used by
// cold code (inversion chain, table construction, point validation); with CAPY_FE_OOL defined the hot
// point arithmetic calls them too, which keeps the scalar-multiplication loops inside the instruction
// cache and the kernels at <= 128 registers (operands then live in local memory / L1).
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void fe_mul_call(Fe* r, const Fe* a, const Fe* b) {
  Fe t;
  fe_mul_inl(t, *a, *b);
  fe_copy(*r, t);
}
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void fe_sqr_call(Fe* r, const Fe* a) {
  Fe t;
  fe_sqr_inl(t, *a);
  fe_copy(*r, t);
}
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void fe_sqrn_call(Fe* r, const Fe* a, int n) {
  Fe t;
  fe_sqr_inl(t, *a);
#pragma unroll 1
  for (int i = 1; i < n; i++) fe_sqr_inl(t, t);
  fe_copy(*r, t);
}

#if defined(CAPY_FE_OOL)
CAPY_HD void fe_mul(Fe& r, const Fe& a, const Fe& b) { fe_mul_call(&r, &a, &b); }
CAPY_HD void fe_sqr(Fe& r, const Fe& a) { fe_sqr_call(&r, &a); }
#else
CAPY_HD void fe_mul(Fe& r, const Fe& a, const Fe& b) { fe_mul_inl(r, a, b); }
CAPY_HD void fe_sqr(Fe& r, const Fe& a) { fe_sqr_inl(r, a); }
#endif

// r = a^(p-2) = 1/a (0 -> 0).  p - 2 = [223 ones][0][222 ones][0][1]  
CAPY_HD void fe_inv(Fe& r, const Fe& x) {
  Fe t3, t, a;
  // t3 = x^(2^3 - 1)
  fe_sqrn_call(&a, &x, 1);     fe_mul_call(&t, &a, &x);
  fe_sqrn_call(&a, &t, 1);     fe_mul_call(&t3, &a, &x);
  fe_sqrn_call(&a, &t3, 3);    fe_mul_call(&t, &a, &t3);    // 2^6 - 1
  Fe u;
  fe_sqrn_call(&a, &t, 6);     fe_mul_call(&u, &a, &t);     // 2^12 - 1
  fe_sqrn_call(&a, &u, 12);    fe_mul_call(&t, &a, &u);     // 2^24 - 1
  fe_sqrn_call(&a, &t, 3);     fe_mul_call(&u, &a, &t3);    // 2^27 - 1
  fe_sqrn_call(&a, &u, 27);    fe_mul_call(&t, &a, &u);     // 2^54 - 1
  fe_sqrn_call(&a, &t, 54);    fe_mul_call(&u, &a, &t);     // 2^108 - 1
  fe_sqrn_call(&a, &u, 3);     fe_mul_call(&t, &a, &t3);    // 2^111 - 1
  fe_sqrn_call(&a, &t, 111);   fe_mul_call(&u, &a, &t);     // 2^222 - 1  (u)
  fe_sqrn_call(&a, &u, 1);     fe_mul_call(&t, &a, &x);     // 2^223 - 1  (t)
  fe_sqrn_call(&a, &t, 223);   fe_mul_call(&a, &a, &u);
  fe_sqrn_call(&a, &a, 2);     fe_mul_call(&r, &a, &x);
}

// fully reduce to the canonical representative in [0, p)
CAPY_HD void fe_canon(Fe& a) {
  fe_weak(a);
  // serial carry until every limb < 2^28 (value then < 2^448 < 2p)
#pragma unroll 1
  for (int pass = 0; pass < 3; pass++) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t v = a.v[i] + c;
      a.v[i] = v & M28;
      c = v >> 28;
    }
    a.v[0] += c;
    a.v[8] += c;
  }
  // a >= p  <=>  a + 2^224 + 1 carries out of 2^448
  uint32_t t[16], c = 1;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t v = a.v[i] + c + (i == 8 ? 1u : 0u);
    t[i] = v & M28;
    c = v >> 28;
  }
  const uint32_t m = 0u - c;  // all ones if a >= p
#pragma unroll
  for (int i = 0; i < 16; i++) a.v[i] = (t[i] & m) | (a.v[i] & ~m);
}

// 56 bytes little-endian (FieldElement::to_bytes) <-> limbs.  `w` = 14 little-endian u32 words.
CAPY_HD void fe_from_words(Fe& r, const uint32_t (&w)[14]) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int bit = 28 * i, q = bit >> 5, sh = bit & 31;
    uint32_t v = w[q] >> sh;
    if (sh > 4 && q + 1 < 14) v |= w[q + 1] << (32 - sh);
    r.v[i] = v & M28;
  }
}
CAPY_HD void fe_to_words(uint32_t (&w)[14], const Fe& a_in) {
  Fe a;
  fe_copy(a, a_in);
  fe_canon(a);
#pragma unroll
  for (int q = 0; q < 14; q++) w[q] = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int bit = 28 * i, q = bit >> 5, sh = bit & 31;
    w[q] |= a.v[i] << sh;
    if (sh > 4 && q + 1 < 14) w[q + 1] |= a.v[i] >> (32 - sh);
  }
}

CAPY_HD bool fe_is_zero(const Fe& a_in) {
  Fe a;
  fe_copy(a, a_in);
  fe_canon(a);
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc |= a.v[i];
  return acc == 0;
}

// constant-time select: r = m ? b : r  (m = all ones / all zeros)
CAPY_HD void fe_cmov(Fe& r, const Fe& b, uint32_t m) {
#pragma unroll
  for (int i = 0; i < 16; i++) r.v[i] = (b.v[i] & m) | (r.v[i] & ~m);
}

}  // namespace capy
