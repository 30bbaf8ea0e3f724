// keccak_pair_probe.cu -- development tool: (1) checks keccak_f1600_pair against keccak_f1600 on the device,
// (2) times a chain of permutations per thread / per thread pair at one warp per scheduler.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "keccak_pair.cuh"
using namespace capy;

__global__ void __launch_bounds__(128) chain_single(uint64_t* st, int nperm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  Lane a[25];
  for (int k = 0; k < 25; k++) { uint64_t v = st[(size_t)t * 25 + k]; a[k].lo = (uint32_t)v; a[k].hi = (uint32_t)(v >> 32); }
  for (int p = 0; p < nperm; p++) keccak_f1600(a);
  for (int k = 0; k < 25; k++) st[(size_t)t * 25 + k] = ((uint64_t)a[k].hi << 32) | a[k].lo;
}
__global__ void __launch_bounds__(128) chain_pair(uint64_t* st, int nperm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = t >> 1;
  const uint32_t half = t & 1;
  uint32_t h[25];
  for (int k = 0; k < 25; k++) { uint64_t v = st[(size_t)item * 25 + k]; h[k] = half ? (uint32_t)(v >> 32) : (uint32_t)v; }
  for (int p = 0; p < nperm; p++) keccak_f1600_pair(h, half);
  uint32_t* o = reinterpret_cast<uint32_t*>(st);
  for (int k = 0; k < 25; k++) o[((size_t)item * 25 + k) * 2 + half] = h[k];
}

// the pair chain with an absorb in front of every permutation: 9 words per thread from a per-pair stream (pairs of a warp
// read 16 different streams, like 16 different messages), next block prefetched -- MODE 0: 32-bit loads of the own halves,
// MODE 1: one 64-bit load per lane by alternating threads + exchange of the foreign halves by shuffle
template <int MODE>
__global__ void __launch_bounds__(128) chain_pair_feed(uint64_t* st, int nperm, const uint32_t* __restrict__ feed, size_t stream_words) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = t >> 1;
  const uint32_t half = t & 1;
  uint32_t h[25];
  for (int k = 0; k < 25; k++) { uint64_t v = st[(size_t)item * 25 + k]; h[k] = half ? (uint32_t)(v >> 32) : (uint32_t)v; }
  const uint32_t* q = feed + (size_t)item * stream_words;
  uint32_t c[9];
  uint2 d[5];
  if (MODE == 0) { for (int j = 0; j < 9; j++) c[j] = __ldg(q + 2 * j + half); }
  else { for (int j = 0; j < 5; j++) if (2 * j + (int)half < 9) d[j] = __ldg(reinterpret_cast<const uint2*>(q) + 2 * j + half); }
  for (int p = 0; p < nperm; p++) {
    if (MODE == 0) {
      for (int j = 0; j < 9; j++) h[j] ^= c[j];
      q += 18;
      for (int j = 0; j < 9; j++) c[j] = __ldg(q + 2 * j + half);
    } else {
      // thread `half` holds lanes 2j + half entirely; it keeps its own half and hands the other half to the partner
      for (int j = 0; j < 5; j++) {
        const uint32_t mine = half ? d[j].y : d[j].x, theirs = half ? d[j].x : d[j].y;
        const uint32_t got = __shfl_xor_sync(0xffffffffu, theirs, 1);  // partner's lane 2j + (1 - half), my half of it
        if (2 * j + (int)half < 9) h[2 * j + half] ^= mine;
        if (2 * j + 1 - (int)half < 9) h[2 * j + 1 - half] ^= got;
      }
      q += 18;
      for (int j = 0; j < 5; j++) if (2 * j + (int)half < 9) d[j] = __ldg(reinterpret_cast<const uint2*>(q) + 2 * j + half);
    }
    keccak_f1600_pair(h, half);
  }
  uint32_t* o = reinterpret_cast<uint32_t*>(st);
  for (int k = 0; k < 25; k++) o[((size_t)item * 25 + k) * 2 + half] = h[k];
}

// one state per warp: thread l < 25 owns lane l = x + 5y
__constant__ uint8_t RHO25[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
__global__ void __launch_bounds__(128) chain_warp25(uint64_t* st, int nperm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = t >> 5;
  const int l = threadIdx.x & 31;
  const int ll = l < 25 ? l : 0;
  const int x = ll % 5, y = ll / 5;
  uint64_t v = st[(size_t)item * 25 + ll];
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  const int c1 = (ll + 5) % 25, c2 = (ll + 10) % 25, c3 = (ll + 15) % 25, c4 = (ll + 20) % 25;
  const int xm = (x + 4) % 5 + 5 * y, xp = (x + 1) % 5 + 5 * y;
  const uint32_t rho = RHO25[ll];
  const uint32_t swap = rho >= 32, m = rho & 31;
  // destination (X', Y') = (x, y): B[X'][Y'] comes from lane ((X' + 3Y') % 5) + 5 X'
  const int s0 = ((x + 3 * y) % 5) + 5 * x;
  const int s1 = (((x + 1) % 5 + 3 * y) % 5) + 5 * ((x + 1) % 5);
  const int s2 = (((x + 2) % 5 + 3 * y) % 5) + 5 * ((x + 2) % 5);
  const uint32_t is0 = ll == 0 && l == 0 ? 0xffffffffu : 0u;
  const unsigned FULL = 0xffffffffu;
  for (int p = 0; p < nperm; p++) {
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
      uint32_t clo = lop_xor3(lop_xor3(lo, __shfl_sync(FULL, lo, c1), __shfl_sync(FULL, lo, c2)), __shfl_sync(FULL, lo, c3), __shfl_sync(FULL, lo, c4));
      uint32_t chi = lop_xor3(lop_xor3(hi, __shfl_sync(FULL, hi, c1), __shfl_sync(FULL, hi, c2)), __shfl_sync(FULL, hi, c3), __shfl_sync(FULL, hi, c4));
      const uint32_t mlo = __shfl_sync(FULL, clo, xm), mhi = __shfl_sync(FULL, chi, xm);
      const uint32_t plo = __shfl_sync(FULL, clo, xp), phi = __shfl_sync(FULL, chi, xp);
      uint32_t tlo = lop_xor3(lo, mlo, __funnelshift_l(phi, plo, 1));
      uint32_t thi = lop_xor3(hi, mhi, __funnelshift_l(plo, phi, 1));
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
  if (l < 25) st[(size_t)item * 25 + l] = ((uint64_t)hi << 32) | lo;
}

// One state per warp again, but the lanes are exchanged through SHARED MEMORY instead of shuffles (an attempt to get under
// the ~180 clocks per round of the shuffle version: 64-bit LDS fetch both halves of a lane at once, so a round issues
// 3 STS.64 + 8 LDS.64 instead of 18 SHFL; the three dependent exchange stages per round remain).
__global__ void __launch_bounds__(128) chain_warp25_smem(uint64_t* st, int nperm) {
  __shared__ uint2 ex[4][3][32];  // per warp: the lanes, the column parities, the rotated lanes
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = t >> 5, w = threadIdx.x >> 5;
  const int l = threadIdx.x & 31;
  const int ll = l < 25 ? l : 0;
  const int x = ll % 5, y = ll / 5;
  uint64_t v = st[(size_t)item * 25 + ll];
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  const int c1 = (ll + 5) % 25, c2 = (ll + 10) % 25, c3 = (ll + 15) % 25, c4 = (ll + 20) % 25;
  const int xm = (x + 4) % 5 + 5 * y, xp = (x + 1) % 5 + 5 * y;
  const uint32_t rho = RHO25[ll];
  const uint32_t swap = rho >= 32, m = rho & 31;
  const int s0 = ((x + 3 * y) % 5) + 5 * x;
  const int s1 = (((x + 1) % 5 + 3 * y) % 5) + 5 * ((x + 1) % 5);
  const int s2 = (((x + 2) % 5 + 3 * y) % 5) + 5 * ((x + 2) % 5);
  const uint32_t is0 = ll == 0 && l == 0 ? 0xffffffffu : 0u;
  uint2* A = ex[w][0];
  uint2* C = ex[w][1];
  uint2* E = ex[w][2];
  for (int p = 0; p < nperm; p++) {
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
      A[l] = make_uint2(lo, hi);
      __syncwarp();
      const uint2 a1 = A[c1], a2 = A[c2], a3 = A[c3], a4 = A[c4];
      const uint32_t clo = lop_xor3(lop_xor3(lo, a1.x, a2.x), a3.x, a4.x);
      const uint32_t chi = lop_xor3(lop_xor3(hi, a1.y, a2.y), a3.y, a4.y);
      C[l] = make_uint2(clo, chi);
      __syncwarp();
      const uint2 cm = C[xm], cp = C[xp];
      const uint32_t tlo = lop_xor3(lo, cm.x, __funnelshift_l(cp.y, cp.x, 1));
      const uint32_t thi = lop_xor3(hi, cm.y, __funnelshift_l(cp.x, cp.y, 1));
      const uint32_t alo = swap ? thi : tlo, ahi = swap ? tlo : thi;
      E[l] = make_uint2(__funnelshift_l(ahi, alo, m), __funnelshift_l(alo, ahi, m));
      __syncwarp();
      const uint2 b0 = E[s0], b1 = E[s1], b2 = E[s2];
      const uint2 rc = KECCAK_RC[r];
      lo = lop_chi(b0.x, b1.x, b2.x) ^ (rc.x & is0);
      hi = lop_chi(b0.y, b1.y, b2.y) ^ (rc.y & is0);
    }
  }
  if (l < 25) st[(size_t)item * 25 + l] = ((uint64_t)hi << 32) | lo;
}

// One SHEET (column x: lanes A[x][0..4]) per thread, five threads per state, six states per warp (lanes 30, 31 idle).
// theta: the column parity is local, D needs the parities of the two neighbour columns (4 SHFL); rho: five local
// rotations by per-thread amounts; pi: B[y][2x+3y] = A[x][y] is a 5 x 5 transposition between the five threads (10 SHFL,
// the source thread picks the row its reader wants after a barrel rotation of its five lanes by x); chi: every row needs
// the lanes of columns x+1 and x+2 (20 SHFL).  34 SHFL + ~80 ALU instructions per round and warp for six states.
__global__ void __launch_bounds__(128) chain_sheet5(uint64_t* st, int nperm, int n_items) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = threadIdx.x & 31;
  const int g = l / 5 < 6 ? l / 5 : 5, x = l < 30 ? l % 5 : 0;
  const bool live = l < 30;
  const int item0 = (t >> 5) * 6 + g;
  const int item = item0 < n_items ? item0 : 0;
  const int base = 5 * g;
  Lane a[5];
#pragma unroll
  for (int y = 0; y < 5; y++) {
    const uint64_t v = st[(size_t)item * 25 + x + 5 * y];
    a[y].lo = (uint32_t)v;
    a[y].hi = (uint32_t)(v >> 32);
  }
  uint32_t rs[5], rm[5];  // rho: rotation by 32 * rs + rm for the lane of row y
#pragma unroll
  for (int y = 0; y < 5; y++) {
    const uint32_t r = RHO25[x + 5 * y];
    rs[y] = r >= 32;
    rm[y] = r & 31;
  }
  const int src_m = base + (x + 4) % 5, src_p = base + (x + 1) % 5, src_p2 = base + (x + 2) % 5;
  int pi_src[5];
#pragma unroll
  for (int Y = 0; Y < 5; Y++) pi_src[Y] = base + (3 * Y + x) % 5;
  const bool b0 = x & 1, b1 = x & 2, b2 = x & 4;
  const uint32_t is0 = (live && x == 0) ? 0xffffffffu : 0u;
  const unsigned FULL = 0xffffffffu;
  for (int p = 0; p < nperm; p++) {
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
      const uint32_t clo = lop_xor3(lop_xor3(a[0].lo, a[1].lo, a[2].lo), a[3].lo, a[4].lo);
      const uint32_t chi = lop_xor3(lop_xor3(a[0].hi, a[1].hi, a[2].hi), a[3].hi, a[4].hi);
      const uint32_t mlo = __shfl_sync(FULL, clo, src_m), mhi = __shfl_sync(FULL, chi, src_m);
      const uint32_t plo = __shfl_sync(FULL, clo, src_p), phi = __shfl_sync(FULL, chi, src_p);
      const uint32_t r1lo = __funnelshift_l(phi, plo, 1), r1hi = __funnelshift_l(plo, phi, 1);
      Lane e[5];
#pragma unroll
      for (int y = 0; y < 5; y++) {
        const uint32_t tlo = lop_xor3(a[y].lo, mlo, r1lo), thi = lop_xor3(a[y].hi, mhi, r1hi);
        const uint32_t ulo = rs[y] ? thi : tlo, uhi = rs[y] ? tlo : thi;
        e[y].lo = __funnelshift_l(uhi, ulo, rm[y]);
        e[y].hi = __funnelshift_l(ulo, uhi, rm[y]);
      }
      // f[j] = e[(j + x) % 5]: barrel rotation of the five lanes by x (1, 2, 4)
      Lane f[5];
#pragma unroll
      for (int j = 0; j < 5; j++) f[j] = b0 ? e[(j + 1) % 5] : e[j];
#pragma unroll
      for (int j = 0; j < 5; j++) e[j] = b1 ? f[(j + 2) % 5] : f[j];
#pragma unroll
      for (int j = 0; j < 5; j++) f[j] = b2 ? e[(j + 4) % 5] : e[j];
      // pi: row Y of the new column comes from thread (3Y + x) % 5, which offers its lane of row (s + 2Y) % 5 = f[2Y % 5]
      Lane b[5];
#pragma unroll
      for (int Y = 0; Y < 5; Y++) {
        b[Y].lo = __shfl_sync(FULL, f[(2 * Y) % 5].lo, pi_src[Y]);
        b[Y].hi = __shfl_sync(FULL, f[(2 * Y) % 5].hi, pi_src[Y]);
      }
      const uint2 rc = KECCAK_RC[r];
#pragma unroll
      for (int Y = 0; Y < 5; Y++) {
        const uint32_t n1l = __shfl_sync(FULL, b[Y].lo, src_p), n1h = __shfl_sync(FULL, b[Y].hi, src_p);
        const uint32_t n2l = __shfl_sync(FULL, b[Y].lo, src_p2), n2h = __shfl_sync(FULL, b[Y].hi, src_p2);
        a[Y].lo = lop_chi(b[Y].lo, n1l, n2l);
        a[Y].hi = lop_chi(b[Y].hi, n1h, n2h);
      }
      a[0].lo ^= rc.x & is0;
      a[0].hi ^= rc.y & is0;
    }
  }
  if (live && item0 < n_items) {
#pragma unroll
    for (int y = 0; y < 5; y++) st[(size_t)item * 25 + x + 5 * y] = ((uint64_t)a[y].hi << 32) | a[y].lo;
  }
}

// the same permutation through the WarpKeccak struct the sponge uses, with an early exit in front (what the
// tiered kernel has) and one XOR per permutation standing in for the absorb
__global__ void __launch_bounds__(128) chain_warp25_struct(uint64_t* st, int nperm, int n_items, const uint32_t* __restrict__ feed) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = t >> 5;
  if (item >= n_items) return;
  const int l = threadIdx.x & 31;
  WarpKeccak wk;
  wk.init(l);
  const uint64_t v = st[(size_t)item * 25 + (l < 25 ? l : 0)];
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  uint32_t w = feed ? __ldg(feed + t) : 0u;
  for (int p = 0; p < nperm; p++) {
    if (l < 9) lo ^= w;
    if (feed && p + 1 < nperm) w = __ldg(feed + t + (size_t)(p + 1) * 32);
    wk.permute(lo, hi);
  }
  if (l < 25) st[(size_t)item * 25 + l] = ((uint64_t)hi << 32) | lo;
}

int main(int argc, char** argv) {
  const int nperm = argc > 1 ? atoi(argv[1]) : 2000;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  for (int wps = 1; wps <= 4; wps *= 2) {  // warps per scheduler
    const int blocks = sms * wps, threads = blocks * 128;
    std::vector<uint64_t> init((size_t)threads * 25);
    uint64_t x = 88172645463325252ull;
    for (auto& v : init) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x; }
    uint64_t *d1, *d2;
    cudaMalloc(&d1, init.size() * 8); cudaMalloc(&d2, init.size() * 8);
    cudaMemcpy(d1, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d2, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms1, ms2;
    chain_single<<<blocks, 128>>>(d1, 2); chain_pair<<<blocks, 128>>>(d2, 2);  // warm-up (also part of the check)
    cudaEventRecord(e0); chain_single<<<blocks, 128>>>(d1, nperm); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms1, e0, e1);
    cudaEventRecord(e0); chain_pair<<<blocks, 128>>>(d2, nperm); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms2, e0, e1);
    {
      uint64_t* d3; cudaMalloc(&d3, init.size() * 8);
      cudaMemcpy(d3, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
      chain_warp25<<<blocks, 128>>>(d3, 2);
      float ms3;
      cudaEventRecord(e0); chain_warp25<<<blocks, 128>>>(d3, nperm); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms3, e0, e1);
      std::vector<uint64_t> r3((size_t)threads * 25), r1b((size_t)threads * 25);
      cudaMemcpy(r3.data(), d3, r3.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(r1b.data(), d1, r1b.size() * 8, cudaMemcpyDeviceToHost);
      size_t bad3 = 0;
      for (size_t i = 0; i < (size_t)threads / 32 * 25; i++) bad3 += r1b[i] != r3[i];
      printf("{\"warps_per_scheduler\": %d, \"warp25_us_per_perm\": %.4f, \"chain_speedup_vs_single\": %.3f, \"warp25_Gperm_s\": %.4f, \"mismatching_lanes\": %zu}\n",
             wps, ms3 * 1e3 / nperm, ms1 / ms3, threads / 32 / (ms3 * 1e-3 / nperm) / 1e9, bad3);
      cudaFree(d3);
    }
    {
      // sheet-per-thread: six states per warp
      const int n_items8 = threads / 32 * 6;
      uint64_t* d8; cudaMalloc(&d8, init.size() * 8);
      cudaMemcpy(d8, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
      chain_sheet5<<<blocks, 128>>>(d8, 2, n_items8);
      float ms8;
      cudaEventRecord(e0); chain_sheet5<<<blocks, 128>>>(d8, nperm, n_items8); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms8, e0, e1);
      std::vector<uint64_t> r8((size_t)threads * 25), r1d((size_t)threads * 25);
      cudaMemcpy(r8.data(), d8, r8.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(r1d.data(), d1, r1d.size() * 8, cudaMemcpyDeviceToHost);
      size_t bad8 = 0;
      for (size_t i = 0; i < (size_t)n_items8 * 25; i++) bad8 += r1d[i] != r8[i];
      printf("{\"warps_per_scheduler\": %d, \"sheet5_six_states_per_warp_us_per_perm\": %.4f, \"states_per_sm\": %d, \"Gperm_s\": %.4f, \"mismatching_lanes\": %zu}\n",
             wps, ms8 * 1e3 / nperm, 24 * wps, n_items8 / (ms8 * 1e-3 / nperm) / 1e9, bad8);
      cudaFree(d8);
    }
    {
      uint64_t* d7; cudaMalloc(&d7, init.size() * 8);
      cudaMemcpy(d7, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
      chain_warp25_smem<<<blocks, 128>>>(d7, 2);
      float ms7;
      cudaEventRecord(e0); chain_warp25_smem<<<blocks, 128>>>(d7, nperm); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms7, e0, e1);
      std::vector<uint64_t> r7((size_t)threads * 25), r1c((size_t)threads * 25);
      cudaMemcpy(r7.data(), d7, r7.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(r1c.data(), d1, r1c.size() * 8, cudaMemcpyDeviceToHost);
      size_t bad7 = 0;
      for (size_t i = 0; i < (size_t)threads / 32 * 25; i++) bad7 += r1c[i] != r7[i];
      printf("{\"warps_per_scheduler\": %d, \"warp25_shared_memory_exchange_us_per_perm\": %.4f, \"mismatching_lanes\": %zu}\n", wps,
             ms7 * 1e3 / nperm, bad7);
      cudaFree(d7);
    }
    {
      uint64_t* d4; cudaMalloc(&d4, init.size() * 8);
      cudaMemcpy(d4, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
      chain_warp25_struct<<<blocks, 128>>>(d4, 2, threads / 32, nullptr);
      float ms4;
      cudaEventRecord(e0); chain_warp25_struct<<<blocks, 128>>>(d4, nperm, threads / 32, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms4, e0, e1);
      uint32_t* feed; cudaMalloc(&feed, (size_t)threads * (nperm + 1) * 4); cudaMemset(feed, 0x5a, (size_t)threads * (nperm + 1) * 4);
      float ms5;
      cudaEventRecord(e0); chain_warp25_struct<<<blocks, 128>>>(d4, nperm, threads / 32, feed); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms5, e0, e1);
      cudaFree(feed);
      printf("{\"warps_per_scheduler\": %d, \"warp25_struct_early_exit_us_per_perm\": %.4f, \"with_one_global_load_per_perm\": %.4f}\n", wps,
             ms4 * 1e3 / nperm, ms5 * 1e3 / nperm);
      cudaFree(d4);
    }
    if (wps == 1) {
      const size_t stream_words = (size_t)18 * (nperm + 2);
      uint32_t* feed; cudaMalloc(&feed, (size_t)(threads / 2) * stream_words * 4); cudaMemset(feed, 0x3c, (size_t)(threads / 2) * stream_words * 4);
      uint64_t *d5, *d6; cudaMalloc(&d5, init.size() * 8); cudaMalloc(&d6, init.size() * 8);
      cudaMemcpy(d5, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(d6, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
      float m0, m1;
      chain_pair_feed<0><<<blocks, 128>>>(d5, 2, feed, stream_words); chain_pair_feed<1><<<blocks, 128>>>(d6, 2, feed, stream_words);
      cudaEventRecord(e0); chain_pair_feed<0><<<blocks, 128>>>(d5, nperm, feed, stream_words); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&m0, e0, e1);
      cudaEventRecord(e0); chain_pair_feed<1><<<blocks, 128>>>(d6, nperm, feed, stream_words); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&m1, e0, e1);
      std::vector<uint64_t> r5(init.size()), r6(init.size());
      cudaMemcpy(r5.data(), d5, r5.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r6.data(), d6, r6.size() * 8, cudaMemcpyDeviceToHost);
      size_t bad56 = 0; for (size_t i = 0; i < (size_t)threads / 2 * 25; i++) bad56 += r5[i] != r6[i];
      printf("{\"pair_with_absorb_us_per_perm\": {\"own_halves_32bit_loads\": %.4f, \"whole_lanes_64bit_loads_plus_exchange\": %.4f}, \"variants_agree\": %s, \"err\": \"%s\"}\n",
             m0 * 1e3 / nperm, m1 * 1e3 / nperm, bad56 ? "false" : "true", cudaGetErrorString(cudaGetLastError()));
      cudaFree(feed); cudaFree(d5); cudaFree(d6);
    }
    // pair kernel covers threads/2 items: compare those
    std::vector<uint64_t> r1((size_t)threads * 25), r2((size_t)threads * 25);
    cudaMemcpy(r1.data(), d1, r1.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r2.data(), d2, r2.size() * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < (size_t)threads / 2 * 25; i++) bad += r1[i] != r2[i];
    printf("{\"warps_per_scheduler\": %d, \"nperm\": %d, \"single_us_per_perm\": %.4f, \"pair_us_per_perm\": %.4f, "
           "\"chain_speedup\": %.3f, \"single_Gperm_s\": %.3f, \"pair_Gperm_s\": %.3f, \"mismatching_lanes\": %zu, \"err\": \"%s\"}\n",
           wps, nperm, ms1 * 1e3 / nperm, ms2 * 1e3 / nperm, ms1 / ms2, threads / (ms1 * 1e-3 / nperm) / 1e9,
           threads / 2 / (ms2 * 1e-3 / nperm) / 1e9, bad, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d1); cudaFree(d2);
  }
  return 0;
}
