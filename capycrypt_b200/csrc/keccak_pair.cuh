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

// both threads of the pair must call this together.  Four rounds per loop iteration: inside the sponge the loop
// counter is not provably warp-uniform, so every iteration ends in a vector branch plus a convergence check in
// front of the shuffles (~40 clocks at one warp per scheduler, where nothing hides them).
__device__ __forceinline__ void keccak_f1600_pair(uint32_t (&h)[25], uint32_t half, unsigned mask = 0xffffffffu) {
#pragma unroll 4
  for (int r = 0; r < 24; r++) {
    const uint2 rc = KECCAK_RC[r];
    keccak_round_pair(h, half ? rc.y : rc.x, mask);
  }
}

// ---- one state per WARP: thread l < 25 owns lane l = x + 5y (threads 25..31 idle but present in the shuffles) ----
// theta needs the four other lanes of the column and the parities of the two neighbouring columns, rho/pi/chi the
// three source lanes of the row after the permutation: 18 shuffles + ~14 ALU instructions per round and thread.
// One permutation per 2.18 us in isolation at one warp per scheduler (4.55 us with one thread per state) for 8x the
// issue slots per state: only for the few longest messages of a chain-bound batch.
struct WarpKeccak {
  int c1, c2, c3, c4;  // the other four lanes of the column
  int xm, xp;          // lanes holding the parity of columns x - 1 and x + 1 (same row)
  int s0, s1, s2;      // pi: sources of B[x][y], B[x+1][y], B[x+2][y]
  uint32_t swap, m;    // rho: rotation by 32 * swap + m
  uint32_t is0;        // all ones in lane 0 (iota)
  __device__ __forceinline__ void init(int l) {
    const int ll = l < 25 ? l : 0, x = ll % 5, y = ll / 5;
    c1 = (ll + 5) % 25; c2 = (ll + 10) % 25; c3 = (ll + 15) % 25; c4 = (ll + 20) % 25;
    xm = (x + 4) % 5 + 5 * y;
    xp = (x + 1) % 5 + 5 * y;
    // B[X][Y] = rot(A'[x][y]) with X = y, Y = 2x + 3y  <=>  the source of (X, Y) is lane ((X + 3Y) % 5) + 5X
    s0 = ((x + 3 * y) % 5) + 5 * x;
    s1 = (((x + 1) % 5 + 3 * y) % 5) + 5 * ((x + 1) % 5);
    s2 = (((x + 2) % 5 + 3 * y) % 5) + 5 * ((x + 2) % 5);
    constexpr uint8_t RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 25; k++) r = (k == ll) ? RHO[k] : r;
    swap = r >= 32;
    m = r & 31;
    is0 = l == 0 ? 0xffffffffu : 0u;
  }
  // all 32 threads of the warp together
  __device__ __forceinline__ void permute(uint32_t& lo, uint32_t& hi) const {
    const unsigned FULL = 0xffffffffu;
    // fully unrolled (24 x ~40 instructions = 16 KB): no loop branch, round constants become immediates
#pragma unroll
    for (int r = 0; r < 24; r++) {
      const uint32_t clo = lop_xor3(lop_xor3(lo, __shfl_sync(FULL, lo, c1), __shfl_sync(FULL, lo, c2)), __shfl_sync(FULL, lo, c3), __shfl_sync(FULL, lo, c4));
      const uint32_t chi = lop_xor3(lop_xor3(hi, __shfl_sync(FULL, hi, c1), __shfl_sync(FULL, hi, c2)), __shfl_sync(FULL, hi, c3), __shfl_sync(FULL, hi, c4));
      const uint32_t mlo = __shfl_sync(FULL, clo, xm), mhi = __shfl_sync(FULL, chi, xm);
      const uint32_t plo = __shfl_sync(FULL, clo, xp), phi = __shfl_sync(FULL, chi, xp);
      const uint32_t tlo = lop_xor3(lo, mlo, __funnelshift_l(phi, plo, 1));
      const uint32_t thi = lop_xor3(hi, mhi, __funnelshift_l(plo, phi, 1));
      const uint32_t alo = swap ? thi : tlo, ahi = swap ? tlo : thi;
      const uint32_t elo = __funnelshift_l(ahi, alo, m), ehi = __funnelshift_l(alo, ahi, m);
      const uint32_t b0l = __shfl_sync(FULL, elo, s0), b0h = __shfl_sync(FULL, ehi, s0);
      const uint32_t b1l = __shfl_sync(FULL, elo, s1), b1h = __shfl_sync(FULL, ehi, s1);
      const uint32_t b2l = __shfl_sync(FULL, elo, s2), b2h = __shfl_sync(FULL, ehi, s2);
      const uint2 rc = KECCAK_RC[r];
      lo = lop_chi(b0l, b1l, b2l) ^ (rc.x & is0);
      hi = lop_chi(b0h, b1h, b2h) ^ (rc.y & is0);
    }
  }
};

}  // namespace capy
