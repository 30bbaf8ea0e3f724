// keccak.cuh -- Keccak-f[1600] for sm_100a, one sponge state per thread, all 25 lanes in
// registers as 50 x u32.  Replaces src/sha3/keccakf.rs:8-423 of the reference for batches.
//
// Instruction budget per round on the 32-bit datapath:
//   theta parity 20 LOP3(xor3) + 10 SHF (rot1 of the parities), theta-apply a ^ C[x-1] ^ rot1(C[x+1]) 50 LOP3,
//   rho 48 SHF (funnel shifts on swapped halves), pi = register renaming, chi 50 LOP3
//   (a ^ (~b & c) is one LOP3, imm 0xD2), iota 2 LOP3  ==> 122 LOP3 + 58 SHF = 180.
// nvcc's default lowering of a 64-bit rotate is IMAD.SHL + SHF.R.U64 + SHF.R.U32.HI + LOP3,
// so rotations are written as explicit __funnelshift_l pairs.
#pragma once
#include <cstdint>

namespace capy {

struct Lane {
  uint32_t lo, hi;
};

static __device__ __constant__ uint2 KECCAK_RC[24] = {
    {0x00000001u, 0x00000000u}, {0x00008082u, 0x00000000u}, {0x0000808Au, 0x80000000u}, {0x80008000u, 0x80000000u},
    {0x0000808Bu, 0x00000000u}, {0x80000001u, 0x00000000u}, {0x80008081u, 0x80000000u}, {0x00008009u, 0x80000000u},
    {0x0000008Au, 0x00000000u}, {0x00000088u, 0x00000000u}, {0x80008009u, 0x00000000u}, {0x8000000Au, 0x00000000u},
    {0x8000808Bu, 0x00000000u}, {0x0000008Bu, 0x80000000u}, {0x00008089u, 0x80000000u}, {0x00008003u, 0x80000000u},
    {0x00008002u, 0x80000000u}, {0x00000080u, 0x80000000u}, {0x0000800Au, 0x00000000u}, {0x8000000Au, 0x80000000u},
    {0x80008081u, 0x80000000u}, {0x00008080u, 0x80000000u}, {0x80000001u, 0x00000000u}, {0x80008008u, 0x80000000u}};

__device__ __forceinline__ uint32_t lop_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// a ^ (~b & c)
__device__ __forceinline__ uint32_t lop_chi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int N>
__device__ __forceinline__ Lane rotl64(Lane a) {
  Lane r;
  if constexpr (N == 0) {
    return a;
  } else if constexpr (N == 32) {
    r.lo = a.hi;
    r.hi = a.lo;
    return r;
  } else if constexpr (N < 32) {
    r.hi = __funnelshift_l(a.lo, a.hi, N);
    r.lo = __funnelshift_l(a.hi, a.lo, N);
    return r;
  } else {
    r.hi = __funnelshift_l(a.hi, a.lo, N - 32);
    r.lo = __funnelshift_l(a.lo, a.hi, N - 32);
    return r;
  }
}

// B[y][2x+3y] = rot(A[x][y] ^ C[x-1] ^ rot1(C[x+1]), rho[x][y]) for lane index x + 5y.  D[x] is never
// formed: the theta "apply" is a single 3-input XOR per half (this is what makes it 122 LOP3 per round).
#define CAPY_RHO_PI(X, Y, R)                                                                  \
  {                                                                                           \
    Lane t;                                                                                   \
    t.lo = lop_xor3(a[(X) + 5 * (Y)].lo, c[((X) + 4) % 5].lo, r1[((X) + 1) % 5].lo);          \
    t.hi = lop_xor3(a[(X) + 5 * (Y)].hi, c[((X) + 4) % 5].hi, r1[((X) + 1) % 5].hi);          \
    b[(Y) + 5 * ((2 * (X) + 3 * (Y)) % 5)] = rotl64<(R)>(t);                                  \
  }

__device__ __forceinline__ void keccak_round(Lane (&a)[25], uint2 rc) {
  Lane c[5], r1[5], b[25];
#pragma unroll
  for (int x = 0; x < 5; x++) {
    c[x].lo = lop_xor3(lop_xor3(a[x].lo, a[x + 5].lo, a[x + 10].lo), a[x + 15].lo, a[x + 20].lo);
    c[x].hi = lop_xor3(lop_xor3(a[x].hi, a[x + 5].hi, a[x + 10].hi), a[x + 15].hi, a[x + 20].hi);
  }
#pragma unroll
  for (int x = 0; x < 5; x++) r1[x] = rotl64<1>(c[x]);
  CAPY_RHO_PI(0, 0, 0)  CAPY_RHO_PI(1, 0, 1)  CAPY_RHO_PI(2, 0, 62) CAPY_RHO_PI(3, 0, 28) CAPY_RHO_PI(4, 0, 27)
  CAPY_RHO_PI(0, 1, 36) CAPY_RHO_PI(1, 1, 44) CAPY_RHO_PI(2, 1, 6)  CAPY_RHO_PI(3, 1, 55) CAPY_RHO_PI(4, 1, 20)
  CAPY_RHO_PI(0, 2, 3)  CAPY_RHO_PI(1, 2, 10) CAPY_RHO_PI(2, 2, 43) CAPY_RHO_PI(3, 2, 25) CAPY_RHO_PI(4, 2, 39)
  CAPY_RHO_PI(0, 3, 41) CAPY_RHO_PI(1, 3, 45) CAPY_RHO_PI(2, 3, 15) CAPY_RHO_PI(3, 3, 21) CAPY_RHO_PI(4, 3, 8)
  CAPY_RHO_PI(0, 4, 18) CAPY_RHO_PI(1, 4, 2)  CAPY_RHO_PI(2, 4, 61) CAPY_RHO_PI(3, 4, 56) CAPY_RHO_PI(4, 4, 14)
#pragma unroll
  for (int y = 0; y < 5; y++) {
#pragma unroll
    for (int x = 0; x < 5; x++) {
      a[x + 5 * y].lo = lop_chi(b[x + 5 * y].lo, b[(x + 1) % 5 + 5 * y].lo, b[(x + 2) % 5 + 5 * y].lo);
      a[x + 5 * y].hi = lop_chi(b[x + 5 * y].hi, b[(x + 1) % 5 + 5 * y].hi, b[(x + 2) % 5 + 5 * y].hi);
    }
  }
  a[0].lo ^= rc.x;
  a[0].hi ^= rc.y;
}
#undef CAPY_RHO_PI

#ifndef CAPY_KECCAK_UNROLL
#define CAPY_KECCAK_UNROLL 1
#endif

template <int UNROLL = CAPY_KECCAK_UNROLL>
__device__ __forceinline__ void keccak_f1600(Lane (&a)[25]) {
#pragma unroll UNROLL
  for (int r = 0; r < 24; r++) keccak_round(a, KECCAK_RC[r]);
}

__device__ __forceinline__ void state_zero(Lane (&a)[25]) {
#pragma unroll
  for (int i = 0; i < 25; i++) a[i].lo = a[i].hi = 0u;
}

}  // namespace capy
