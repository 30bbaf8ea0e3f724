// keccak_pair.cuh -- Keccak-f[1600] with ONE sponge state shared by TWO adjacent threads of a warp: the even
// thread owns the low 32 bits of every lane, the odd thread the high 32 bits (same algorithm as
// src/sha3/keccakf.rs:8-423, different data distribution).
//
// Why: a sponge is sequential per message, so a batch cannot finish before its longest message has (BASELINE
// config 5: 1 MiB = 14 564 permutations = 68 ms on one thread, because a warp instruction occupies the ALU pipe
// for two clocks however few lanes are active).  Splitting the lanes halves the per-thread work of a round: theta
// parity, theta apply and chi are bitwise and stay local (60 LOP3); a 64-bit rotation by n is ONE funnel shift
// over (own half, partner half) instead of two, with the partner half fetched by a shuffle (29 SHFL + 29 SHF).
// 91 ALU + 29 SHFL instructions per thread and round instead of 180: the chain of one message advances ~1.6x
// faster for ~1.35x the issue slots, which is spent only on the longest messages of a ragged batch.
#pragma once
#include "keccak.cuh"

namespace capy {

// new own half of rot64(lane, N) from the own and the partner half of the lane
template <int N>
__device__ __forceinline__ uint32_t rot_half(uint32_t own, uint32_t partner) {
  if constexpr (N == 0) return own;
  else if constexpr (N == 32) return partner;
  else if constexpr (N < 32) return __funnelshift_l(partner, own, N);
  else return __funnelshift_l(own, partner, N - 32);
}

#define CAPY_PAIR_RHO_PI(X, Y, R)                                                                  \
  {                                                                                                \
    const uint32_t t = lop_xor3(h[(X) + 5 * (Y)], c[((X) + 4) % 5], r1[((X) + 1) % 5]);            \
    uint32_t p = 0;                                                                                \
    if ((R) != 0) p = __shfl_xor_sync(mask, t, 1);                                                 \
    b[(Y) + 5 * ((2 * (X) + 3 * (Y)) % 5)] = rot_half<(R)>(t, p);                                  \
  }

// rc_half = this thread's half of the round constant; mask = the two lanes of the pair (pairs of one warp may
// sit in different loop iterations, so every shuffle names only its own pair)
__device__ __forceinline__ void keccak_round_pair(uint32_t (&h)[25], uint32_t rc_half, unsigned mask) {
  uint32_t c[5], r1[5], b[25];
#pragma unroll
  for (int x = 0; x < 5; x++) c[x] = lop_xor3(lop_xor3(h[x], h[x + 5], h[x + 10]), h[x + 15], h[x + 20]);
#pragma unroll
  for (int x = 0; x < 5; x++) r1[x] = rot_half<1>(c[x], __shfl_xor_sync(mask, c[x], 1));
  CAPY_PAIR_RHO_PI(0, 0, 0)  CAPY_PAIR_RHO_PI(1, 0, 1)  CAPY_PAIR_RHO_PI(2, 0, 62) CAPY_PAIR_RHO_PI(3, 0, 28) CAPY_PAIR_RHO_PI(4, 0, 27)
  CAPY_PAIR_RHO_PI(0, 1, 36) CAPY_PAIR_RHO_PI(1, 1, 44) CAPY_PAIR_RHO_PI(2, 1, 6)  CAPY_PAIR_RHO_PI(3, 1, 55) CAPY_PAIR_RHO_PI(4, 1, 20)
  CAPY_PAIR_RHO_PI(0, 2, 3)  CAPY_PAIR_RHO_PI(1, 2, 10) CAPY_PAIR_RHO_PI(2, 2, 43) CAPY_PAIR_RHO_PI(3, 2, 25) CAPY_PAIR_RHO_PI(4, 2, 39)
  CAPY_PAIR_RHO_PI(0, 3, 41) CAPY_PAIR_RHO_PI(1, 3, 45) CAPY_PAIR_RHO_PI(2, 3, 15) CAPY_PAIR_RHO_PI(3, 3, 21) CAPY_PAIR_RHO_PI(4, 3, 8)
  CAPY_PAIR_RHO_PI(0, 4, 18) CAPY_PAIR_RHO_PI(1, 4, 2)  CAPY_PAIR_RHO_PI(2, 4, 61) CAPY_PAIR_RHO_PI(3, 4, 56) CAPY_PAIR_RHO_PI(4, 4, 14)
#pragma unroll
  for (int y = 0; y < 5; y++) {
#pragma unroll
    for (int x = 0; x < 5; x++) h[x + 5 * y] = lop_chi(b[x + 5 * y], b[(x + 1) % 5 + 5 * y], b[(x + 2) % 5 + 5 * y]);
  }
  h[0] ^= rc_half;
}
#undef CAPY_PAIR_RHO_PI

// both threads of the pair must call this together
__device__ __forceinline__ void keccak_f1600_pair(uint32_t (&h)[25], uint32_t half, unsigned mask = 0xffffffffu) {
#pragma unroll 1
  for (int r = 0; r < 24; r++) {
    const uint2 rc = KECCAK_RC[r];
    keccak_round_pair(h, half ? rc.y : rc.x, mask);
  }
}

}  // namespace capy
