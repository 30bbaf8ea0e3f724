// sc448.cuh -- integers mod r (the Ed448 prime subgroup order) for sm_100a.
//
// Replaces the `Scalar` arithmetic the reference takes from tiny_ed448_goldilocks 0.1.8 /
// crypto-bigint U448: `bytes_to_scalar` (sha3/aux_functions.rs:102-106, 56 bytes big-endian, not
// reduced), `.mul_mod(&Scalar::from(4))` (ecc/keypair.rs:43, ecc/signable.rs:42), `Scalar * Scalar`
// (ecc/signable.rs:46), `k - h.mul_mod(&s)` (ecc/signable.rs:54), `scalar_to_bytes` (:108-110).
// Canonical semantics: every result is the representative in [0, r).
// 14 x 32-bit little-endian limbs; this is a few hundred instructions per signature, so clarity wins.
#pragma once
#include <cstdint>

#include "fp448.cuh"

namespace capy {

struct Sc {
  uint32_t w[14];
};

// r = 2^446 - c
#define CAPY_R_LIMBS                                                                                   \
  {0xab5844f3u, 0x2378c292u, 0x8dc58f55u, 0x216cc272u, 0xaed63690u, 0xc44edb49u, 0x7cca23e9u,          \
   0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu}
#define CAPY_C_LIMBS {0x54a7bb0du, 0xdc873d6du, 0x723a70aau, 0xde933d8du, 0x5129c96fu, 0x3bb124b6u, 0x8335dc16u}

// 56 bytes big-endian at p -> limbs (no reduction)
CAPY_HD void sc_from_be(Sc& s, const uint8_t* p) {
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const uint8_t* q = p + 52 - 4 * i;
    s.w[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
}
CAPY_HD void sc_to_be(uint8_t* p, const Sc& s) {
#pragma unroll
  for (int i = 0; i < 14; i++) {
    uint8_t* q = p + 52 - 4 * i;
    q[0] = (uint8_t)(s.w[i] >> 24);
    q[1] = (uint8_t)(s.w[i] >> 16);
    q[2] = (uint8_t)(s.w[i] >> 8);
    q[3] = (uint8_t)s.w[i];
  }
}

// one fold at bit 446: x (N limbs) = hi * 2^446 + lo  ->  y = lo + hi * c  (== x mod r), M limbs.
// All loops have compile-time bounds, so after unrolling every index is static (registers, no local
// memory).  M is chosen by the caller so that the result fits: M = max(14, N - 6).
template <int N, int M>
CAPY_HD void sc_fold(uint32_t (&y)[M], const uint32_t (&x)[N]) {
  const uint32_t CL[7] = CAPY_C_LIMBS;
  constexpr int HL = N - 13;  // limbs of hi (bit 446 = limb 13, bit 30)
  uint32_t hi[HL];
#pragma unroll
  for (int i = 0; i < HL; i++) hi[i] = (x[13 + i] >> 30) | (13 + i + 1 < N ? (x[14 + i] << 2) : 0u);
#pragma unroll
  for (int i = 0; i < M; i++) y[i] = i < 13 ? x[i] : 0u;
  y[13] = x[13] & 0x3fffffffu;
#pragma unroll
  for (int i = 0; i < HL; i++) {
    uint64_t cy = 0;
#pragma unroll
    for (int j = 0; j < 7; j++) {
      if (i + j < M) {
        const uint64_t t = (uint64_t)hi[i] * CL[j] + y[i + j] + cy;
        y[i + j] = (uint32_t)t;
        cy = t >> 32;
      }
    }
#pragma unroll
    for (int k = i + 7; k < M; k++) {
      const uint64_t t = (uint64_t)y[k] + cy;
      y[k] = (uint32_t)t;
      cy = t >> 32;
    }
  }
}

// final step: v < 2^446 + small (14 limbs, top two bits of limb 13 may hold a tiny overflow is NOT allowed:
// callers fold until v < 2^447) -> v mod r with conditional subtractions
CAPY_HD void sc_final(Sc& o, const uint32_t (&v)[14]) {
  const uint32_t RL[14] = CAPY_R_LIMBS;
  uint32_t cur[14];
#pragma unroll
  for (int i = 0; i < 14; i++) cur[i] = v[i];
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {  // v < 2^447 < 3r: two conditional subtractions suffice
    uint32_t t[14];
    uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
      const uint64_t d = (uint64_t)cur[i] - RL[i] - bw;
      t[i] = (uint32_t)d;
      bw = (d >> 63) & 1u;
    }
    const uint32_t keep = 0u - (uint32_t)bw;  // borrow -> cur < r -> keep cur
#pragma unroll
    for (int i = 0; i < 14; i++) cur[i] = (cur[i] & keep) | (t[i] & ~keep);
  }
#pragma unroll
  for (int i = 0; i < 14; i++) o.w[i] = cur[i];
}

// x: N limbs -> x mod r.  Fold sizes: 28 -> 22 -> 16 -> 14 -> 14 (each fold removes ~222 bits:
// c < 2^224, so hi * c has (32 N - 446) + 224 bits).
template <int N>
CAPY_HD void sc_reduce(Sc& o, const uint32_t (&x)[N]) {
  if constexpr (N > 14) {
    constexpr int M = (N - 6) > 14 ? (N - 6) : 14;
    uint32_t y[M];
    sc_fold<N, M>(y, x);
    sc_reduce<M>(o, y);
  } else {
    // N == 14: x < 2^448.  One fold brings it below 2^446 + 3c < 2^447, then subtract.
    uint32_t y[14];
    sc_fold<14, 14>(y, x);
    sc_final(o, y);
  }
}

CAPY_HD void sc_reduce_448(Sc& o, const Sc& a) {
  uint32_t x[14];
#pragma unroll
  for (int i = 0; i < 14; i++) x[i] = a.w[i];
  sc_reduce<14>(o, x);
}

// o = a * b mod r  (Scalar::mul_mod and `Scalar * Scalar`; a, b any 448-bit integers)
CAPY_HD void sc_mul_mod(Sc& o, const Sc& a, const Sc& b) {
  uint32_t p[28];
#pragma unroll
  for (int i = 0; i < 28; i++) p[i] = 0;
#pragma unroll
  for (int i = 0; i < 14; i++) {
    uint64_t cy = 0;
#pragma unroll
    for (int j = 0; j < 14; j++) {
      const uint64_t t = (uint64_t)a.w[i] * b.w[j] + p[i + j] + cy;
      p[i + j] = (uint32_t)t;
      cy = t >> 32;
    }
    p[i + 14] = (uint32_t)cy;
  }
  sc_reduce<28>(o, p);
}

// o = 4 * a mod r  (.mul_mod(&Scalar::from(4_u64)), a any 448-bit integer)
CAPY_HD void sc_mul4_mod(Sc& o, const Sc& a) {
  uint32_t x[15];
  x[0] = a.w[0] << 2;
#pragma unroll
  for (int i = 1; i < 14; i++) x[i] = (a.w[i] << 2) | (a.w[i - 1] >> 30);
  x[14] = a.w[13] >> 30;
  sc_reduce<15>(o, x);
}

// o = a - b mod r for a, b in [0, r)
CAPY_HD void sc_sub_mod(Sc& o, const Sc& a, const Sc& b) {
  const uint32_t RL[14] = CAPY_R_LIMBS;
  uint32_t t[14];
  uint64_t bw = 0;
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const uint64_t v = (uint64_t)a.w[i] - b.w[i] - bw;
    t[i] = (uint32_t)v;
    bw = (v >> 63) & 1u;
  }
  const uint32_t addr = 0u - (uint32_t)bw;
  uint64_t cy = 0;
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const uint64_t v = (uint64_t)t[i] + (RL[i] & addr) + cy;
    o.w[i] = (uint32_t)v;
    cy = v >> 32;
  }
}

// (r + 1) / 4 = 4^-1 mod r   (r == 3 mod 4)
#define CAPY_INV4_LIMBS                                                                                \
  {0xaad6113du, 0x48de30a4u, 0xa37163d5u, 0x085b309cu, 0x6bb58da4u, 0x7113b6d2u, 0xdf3288fau,          \
   0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x0fffffffu}

}  // namespace capy
