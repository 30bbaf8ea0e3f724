// sha3_api.cu -- SHA3-d / cSHAKE / KMACXOF batch entry points (C ABI in include/capy_gpu.h).
//
// Replaces, for batches, the reference call stack
//   hashable.rs:19-35 -> shake_functions.rs:24-89 -> sponge.rs:10-34 -> keccakf.rs:8-423.
// The SP 800-185 encoders (aux_functions.rs:11-68) run on the host only to build the constant
// prefix block of cSHAKE/KMAC; every byte that is absorbed per item is produced on the device.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "internal.h"
#include "hostbatch.h"
#include "sponge.cuh"

namespace capy {

// ------------------------------------------------------------------------------------------------
// uniform-length SHA3-d kernel: every message has the same length, 8-byte aligned starts.
// All control flow depends only on (len, LANES) so it is warp-uniform; the message bytes go
// straight from global memory into the state registers.  This is the cfg-1 hot kernel
// (2^20 x 64 B: one permutation, 4 x 16 B loads, 2 x 16 B stores per thread).
// ------------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(128)
    sha3_uniform_kernel(const uint8_t* __restrict__ data, uint64_t stride, uint64_t len, uint32_t suffix,
                        uint8_t* __restrict__ out, uint32_t out_bytes, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr uint64_t RATE = 8ull * LANES;
  const uint2* q = reinterpret_cast<const uint2*>(data + i * stride);
  Lane a[25];
  state_zero(a);
  const uint64_t nfull = len / RATE;
  for (uint64_t b = 0; b < nfull; b++) {
#pragma unroll
    for (int j = 0; j < LANES; j++) {
      uint2 v = __ldg(q + j);
      a[j].lo ^= v.x;
      a[j].hi ^= v.y;
    }
    q += LANES;
    keccak_f1600(a);
  }
  // final block: rem message bytes, the suffix byte, then (only if rem + 1 < RATE) zeros..0x80
  const uint32_t rem = (uint32_t)(len - nfull * RATE);
#pragma unroll
  for (int j = 0; j < LANES; j++) {
    const uint32_t o = 8u * j;
    uint32_t lo = 0, hi = 0;
    if (o < rem) {  // the aligned 8-byte word holds at least one message byte
      uint2 v = __ldg(q + j);
      lo = v.x;
      hi = v.y;
      if (o + 8 > rem) {  // partial lane: keep rem - o bytes
        const uint32_t keep = rem - o;  // 1..7
        if (keep < 4) {
          lo &= (1u << (8 * keep)) - 1u;
          hi = 0;
        } else if (keep > 4) {
          hi &= (1u << (8 * (keep - 4))) - 1u;
        } else {
          hi = 0;
        }
      }
    }
    if (rem >= o && rem < o + 8) {
      const uint32_t k = rem - o;
      if (k < 4) lo |= suffix << (8 * k);
      else hi |= suffix << (8 * (k - 4));
    }
    a[j].lo ^= lo;
    a[j].hi ^= hi;
  }
  if (rem + 1 < RATE) a[LANES - 1].hi ^= 0x80000000u;  // quirk Q1: no 0x80 when already aligned
  keccak_f1600(a);

  uint8_t* o = out + i * (uint64_t)out_bytes;
  if ((out_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (16u * j < out_bytes)
        reinterpret_cast<uint4*>(o)[j] = make_uint4(a[2 * j].lo, a[2 * j].hi, a[2 * j + 1].lo, a[2 * j + 1].hi);
  } else {  // 28- and 48-byte digests: 4-byte granules
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (8u * j < out_bytes) reinterpret_cast<uint32_t*>(o)[2 * j] = a[j].lo;
      if (8u * j + 4 < out_bytes) reinterpret_cast<uint32_t*>(o)[2 * j + 1] = a[j].hi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// single-block SHA3-d of messages of exactly 8 * MSG_LANES bytes (MSG_LANES < LANES), 16-byte aligned
// rows: the BASELINE cfg-1 shape (2^20 x 64 B).  Everything about the block layout is a compile-time
// constant: message lanes come in with 128-bit loads, the suffix / 0x80 lanes are immediates, round 0 is
// peeled so that the zero lanes fold away, and the remaining 23 rounds run two per loop iteration.
// ------------------------------------------------------------------------------------------------
template <int LANES, int MSG_LANES>
__global__ void __launch_bounds__(128)
    sha3_short_kernel(const uint4* __restrict__ data, uint64_t stride16, uint32_t suffix, uint4* __restrict__ out,
                      uint32_t out_bytes, uint64_t n) {
  static_assert(MSG_LANES % 2 == 0 && MSG_LANES < LANES, "whole 16-byte loads, one block");
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* q = data + i * stride16;
  Lane a[25];
  state_zero(a);
#pragma unroll
  for (int j = 0; j < MSG_LANES / 2; j++) {
    const uint4 v = __ldg(q + j);
    a[2 * j].lo = v.x;
    a[2 * j].hi = v.y;
    a[2 * j + 1].lo = v.z;
    a[2 * j + 1].hi = v.w;
  }
  a[MSG_LANES].lo = suffix;  // 0x06 (0x86 can only occur for 135-byte messages, which are not 8-byte multiples)
  a[LANES - 1].hi ^= 0x80000000u;  // 8 * MSG_LANES + 1 < rate always holds here: zeros then 0x80 (sponge.rs:89-95)
  keccak_round(a, KECCAK_RC[0]);
#pragma unroll 1
  for (int r = 1; r < 23; r += 2) {
    keccak_round(a, KECCAK_RC[r]);
    keccak_round(a, KECCAK_RC[r + 1]);
  }
  keccak_round(a, KECCAK_RC[23]);
  uint4* o = out + i * (uint64_t)(out_bytes / 16);
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (16u * j < out_bytes) o[j] = make_uint4(a[2 * j].lo, a[2 * j].hi, a[2 * j + 1].lo, a[2 * j + 1].hi);
}

// one thread absorbs the whole-block part of a constant prefix into a cached state
template <int LANES>
__global__ void prefix_state_kernel(const uint8_t* prefix, uint32_t nblocks, uint64_t* state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Lane a[25];
  state_zero(a);
  for (uint32_t b = 0; b < nblocks; b++) {
    const uint8_t* p = prefix + (size_t)b * 8 * LANES;
#pragma unroll
    for (int j = 0; j < LANES; j++) {
      uint64_t v = 0;
      for (int k = 0; k < 8; k++) v |= (uint64_t)p[8 * j + k] << (8 * k);
      a[j].lo ^= (uint32_t)v;
      a[j].hi ^= (uint32_t)(v >> 32);
    }
    keccak_f1600(a);
  }
  for (int k = 0; k < 25; k++) state[k] = ((uint64_t)a[k].hi << 32) | a[k].lo;
}

static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// ------------------------------------------------------------------------------------------------
// Mixed-length batches (BASELINE config 5): longest-processing-time-first order.
// A sponge is sequential per message, so with one thread per message the job cannot finish before
// the longest message does.  Items are bucketed by block count (counting sort on the device) and
// processed longest first; when the longest chain would dominate, the launch is limited to 1-3 warps
// per scheduler (dynamic shared memory as an occupancy throttle) so that chain advances at full speed.
// ------------------------------------------------------------------------------------------------
// 16 384 bins of one block each; longer messages (>= 1.1 MB at the SHA3-512 rate) share the last bin
constexpr uint32_t kLenBins = 1u << 14;
constexpr uint64_t kHostPlanMaxItems = 1ull << 18;  // plan_ragged: largest batch whose histogram is taken on the host

__device__ __forceinline__ uint32_t len_bin(const uint64_t* off, uint64_t i, uint32_t stride_bytes) {
  const uint64_t blocks = (off[i + 1] - off[i]) / stride_bytes;
  return blocks < kLenBins - 1 ? (uint32_t)blocks : kLenBins - 1;
}

__global__ void len_hist_kernel(const uint64_t* __restrict__ off, uint64_t n, uint32_t stride_bytes, uint32_t* hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[len_bin(off, i, stride_bytes)], 1u);
}

// hist[k] := number of items in bins > k (start of bin k in descending order); summary = {total blocks
// (lo, hi), max bin, number of non-empty bins}.  1024 threads x 16 consecutive bins (four 128-bit loads each),
// thread 0 owns the HIGHEST bins.
__global__ void __launch_bounds__(1024) len_scan_kernel(uint32_t* hist, uint32_t* summary) {
  __shared__ unsigned long long tot_s;
  __shared__ uint32_t max_s, bins_s;
  const uint32_t t = threadIdx.x;
  constexpr uint32_t PER = kLenBins / 1024;  // 16
  static_assert(PER == 16, "four uint4 per thread");
  if (t == 0) { tot_s = 0; max_s = 0; bins_s = 0; }
  __syncthreads();
  const uint32_t lo = kLenBins - (t + 1) * PER;  // this thread's bins lo .. lo + 15
  uint32_t c[PER];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint4 v = reinterpret_cast<const uint4*>(hist + lo)[q];
    c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
  }
  uint32_t sum = 0, nb = 0, mx = 0;
  unsigned long long tot = 0;
#pragma unroll
  for (int k = 0; k < (int)PER; k++) {
    sum += c[k];
    if (c[k]) { nb++; mx = lo + k; }  // ascending k: the last non-empty one is the largest
    tot += (unsigned long long)c[k] * (lo + k + 1);
  }
  if (sum) {
    atomicAdd(&tot_s, tot);
    atomicMax(&max_s, mx);
    atomicAdd(&bins_s, nb);
  }
  __syncthreads();
  // inclusive scan of the per-thread sums over the threads (thread 0 = highest bins): warp shuffles, then the 32 warp totals
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((t & 31) >= (uint32_t)d) incl += v;
  }
  __shared__ uint32_t wsum[32];
  if ((t & 31) == 31) wsum[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    uint32_t w = wsum[t];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
      if (t >= (uint32_t)d) w += v;
    }
    wsum[t] = w;
  }
  __syncthreads();
  uint32_t run = incl - sum + ((t >> 5) ? wsum[(t >> 5) - 1] : 0u);  // items in bins above this thread's range
  // descending inside the thread: bin lo + 15 first
#pragma unroll
  for (int k = (int)PER - 1; k >= 0; k--) {
    const uint32_t cnt = c[k];
    c[k] = run;
    run += cnt;
  }
#pragma unroll
  for (int q = 0; q < 4; q++) reinterpret_cast<uint4*>(hist + lo)[q] = make_uint4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
  if (t == 0) {
    summary[0] = (uint32_t)tot_s;
    summary[1] = (uint32_t)(tot_s >> 32);
    summary[2] = max_s;
    summary[3] = bins_s;
  }
}

__global__ void len_scatter_kernel(const uint64_t* __restrict__ off, uint64_t n, uint32_t stride_bytes, uint32_t* cursor,
                                   uint32_t* __restrict__ order) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) order[atomicAdd(&cursor[len_bin(off, i, stride_bytes)], 1u)] = (uint32_t)i;
}

// Measured on B200, per permutation of ONE chain inside the sponge (tools/bench_tier_probe.py, 1 MiB SHA3-512 messages,
// profiles/README.md round 2): one thread per state 4.6 us; a thread pair 3.08 us; a whole warp per state 2.04 us on an
// otherwise idle GPU and 2.06 us with every SM busy.  A warp-tier chain issues only ~32 instructions per ~170-clock round,
// so several can share a scheduler -- but the 18 shuffles of its round go through one shuffle unit per SM (one warp-wide
// shuffle per clock): with c chains per scheduler a round cannot be shorter than 72 c clocks.  Measured: 2.06 / 2.36 /
// 3.0 us per permutation at c = 1 / 2 / 3.
constexpr double kPairChainRatio = 3.08 / 4.6;
constexpr int kMaxWarpCosched = 3;
constexpr double kWarpChainRatio[kMaxWarpCosched + 1] = {0.0, 2.06 / 4.6, 2.36 / 4.6, 3.00 / 4.6};

// Tiers of a chain-bound batch.  cum[k] = number of items in length bins > k (bin = whole blocks of the message).
// Times are in units of one thread-per-state permutation; an SM hosts ONE block: 4 c warp-tier items (c = 1, 2 or 3
// chains per scheduler), 64 pair-tier or 128 thread-tier items.  For a step time T every item takes the cheapest
// class its chain still fits: warp tier with a scheduler to itself for the longest, then two and three chains per
// scheduler, then the pair tier, the rest one thread per item.  The smallest T is taken for which (a) the blocks of the
// fast tiers are all resident from the start and (b) the SM time of all tiers fits 148 T.
static void plan_tiers(const std::vector<uint32_t>& cum, uint64_t n, uint32_t max_blocks, double total_blocks, int sm_count,
                       uint64_t warp_items[kMaxWarpCosched], uint64_t* pair_items, int force_c = 0) {
  for (int c = 0; c < kMaxWarpCosched; c++) warp_items[c] = 0;
  *pair_items = 0;
  const size_t nb = cum.size();
  std::vector<double> work_above(nb);  // blocks in bins > k
  double acc = 0;
  for (size_t k = nb; k-- > 0;) {
    work_above[k] = acc;
    const uint64_t cnt = (k == 0 ? n : cum[k - 1]) - cum[k];
    acc += (double)cnt * (double)(k + 1);
  }
  auto longer_than = [&](double thr, uint64_t* count, double* work) {  // items whose chain (bin + 1) exceeds thr
    if (thr < 1.0) {
      *count = n;
      *work = total_blocks;
      return;
    }
    const size_t k = std::min<size_t>((size_t)thr - 1, nb - 1);
    *count = cum[k];
    *work = work_above[k];
  };
  const double l1 = (double)max_blocks;
  const int c_first = force_c >= 1 && force_c <= kMaxWarpCosched ? force_c : 1;
  for (double T = l1 * kWarpChainRatio[c_first]; T < l1; T *= 1.01) {
    // class boundaries: an item with chain L runs c chains per scheduler iff L * ratio[c] <= T < L * ratio[c + 1]
    uint64_t k_above[kMaxWarpCosched + 2];  // [c] = items too long for warp class c + 1 (and every cheaper class)
    double w_above[kMaxWarpCosched + 2];
    k_above[0] = 0;
    w_above[0] = 0;
    for (int c = 1; c <= kMaxWarpCosched; c++) {
      const int next = c + 1;
      const double ratio = next <= kMaxWarpCosched ? kWarpChainRatio[next] : kPairChainRatio;
      longer_than(T / ratio, &k_above[c], &w_above[c]);
      if (force_c && c != force_c) {  // diagnostics: one warp class only
        k_above[c] = c < force_c ? 0 : k_above[force_c];
        w_above[c] = c < force_c ? 0 : w_above[force_c];
      }
    }
    uint64_t k_fast;
    double w_fast;
    longer_than(T, &k_fast, &w_fast);  // must not run one thread per item
    if (k_above[kMaxWarpCosched] == 0 && T < l1 * kPairChainRatio) continue;  // the pair tier alone cannot finish the longest in T
    uint64_t blocks = 0;
    double sm_time = 0;  // in units of (thread-tier permutation x SM)
    uint64_t cnt[kMaxWarpCosched];
    for (int c = 1; c <= kMaxWarpCosched; c++) {
      cnt[c - 1] = k_above[c] - k_above[c - 1];
      blocks += (cnt[c - 1] + 4 * c - 1) / (4 * c);
      sm_time += (w_above[c] - w_above[c - 1]) * kWarpChainRatio[c] / (4.0 * c);
    }
    const uint64_t k_p = k_fast - k_above[kMaxWarpCosched];
    blocks += (k_p + 63) / 64;
    sm_time += (w_fast - w_above[kMaxWarpCosched]) * kPairChainRatio / 64.0 + (total_blocks - w_fast) / 128.0;
    if (blocks + 1 > (uint64_t)sm_count) continue;
    if (sm_time > 0.95 * T * (double)sm_count) continue;
    // the thread-per-item tier gets the SMs the fast tiers leave (those stay busy for about T: their items are the
    // longest); its items are dispatched longest first, so it needs its share of SMs from the start
    if ((total_blocks - w_fast) / 128.0 > T * (double)((uint64_t)sm_count - blocks)) continue;
    for (int c = 0; c < kMaxWarpCosched; c++) warp_items[c] = cnt[c];
    *pair_items = k_p;
    return;
  }
}

void plan_cache_clear(DeviceCtx& dc);  // ctx.cu

// builds the descending-length order for a ragged batch; decides the occupancy throttle and how many of the
// longest items go to the warp-per-item and the two-threads-per-item tiers.
// The histogram comes back to the host (two small D2H copies + stream synchronisations) because the grid shape
// depends on it.  With the plan cache on (capy_gpu_set_plan_cache) the plan of an offsets array is kept, keyed by
// (device pointer, n, unit): every later call with the same array launches without touching the host, so the
// `_dev` entry points are fully asynchronous in steady state.
static int plan_ragged(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t st, const uint64_t* d_off, uint64_t n,
                       uint32_t stride_bytes, LaunchPlan* plan, bool allow_pair = true, int force_c = 0) {
  *plan = LaunchPlan();
  if (n < 1 || n > 0xffffffffull) return CAPY_OK;
  // Host-buffer entry points: the offsets were staged from a host array that is still alive (stage_packed registers it).
  // The lengths are read there, so nothing comes back from the device and the chunk pipeline of the host call never
  // waits for a stream.  Uniform lengths (the bulk case: fixed-size records) need no device work at all.
  const uint64_t* h_off = host_off_find(dc, d_off, n);
  std::vector<uint32_t> h_cnt;  // host histogram (non-uniform batches up to kHostPlanMaxItems)
  uint64_t h_total = 0;
  uint32_t h_max = 0, h_bins = 0;
  if (h_off) {
    const uint64_t len0 = h_off[1] - h_off[0];
    bool uniform = true;
    for (uint64_t i = 1; i < n && uniform; i++) uniform = (h_off[i + 1] - h_off[i]) == len0;
    if (uniform) {  // nothing to order, no chain stands out
      plan->uniform_blocks = len0 / stride_bytes + 1;
      return CAPY_OK;
    }
    if (n > kHostPlanMaxItems) {
      h_off = nullptr;  // a histogram of millions of items is cheaper on the device, round trip included
    } else {
      h_cnt.assign(kLenBins, 0);
      const double inv = 1.0 / (double)stride_bytes;
      for (uint64_t i = 0; i < n; i++) {
        const uint64_t len = h_off[i + 1] - h_off[i];
        uint64_t q = (uint64_t)((double)len * inv);  // len / stride_bytes without a 64-bit division per item
        if (q * stride_bytes > len) q--;
        else if ((q + 1) * stride_bytes <= len) q++;
        const uint32_t b = q < kLenBins - 1 ? (uint32_t)q : kLenBins - 1;
        if (h_cnt[b]++ == 0) h_bins++;
        if (b > h_max) h_max = b;
        h_total += (uint64_t)b + 1;
      }
    }
  }
  const bool cached = ctx->plan_cache.load() != 0 && force_c == 0 && !h_off;
  if (cached) {
    for (PlanEntry& e : dc.plans)
      if (e.off == d_off && e.n == n && e.unit == stride_bytes && e.allow_pair == allow_pair) {
        e.last_use = ++dc.use_clock;
        *plan = e.plan;
        return CAPY_OK;
      }
  }
  int si = 0;  // scratch pair per internal stream (chunks on different streams overlap); callers' streams use pair 0
  for (int k = 0; k < kNumStreams; k++)
    if (dc.streams[k] == st) si = k;
  uint32_t* hist = (uint32_t*)scratch_get(dc, 18 + 2 * si, (size_t)kLenBins * 4 + 64);
  uint32_t* order = nullptr;
  uint32_t* owned = nullptr;
  if (cached) {
    if (cudaMalloc(&owned, (size_t)n * 4) != cudaSuccess) {
      cudaGetLastError();
      return CAPY_ERR_OOM;
    }
    order = owned;
  } else {
    order = (uint32_t*)scratch_get(dc, 19 + 2 * si, (size_t)n * 4);
  }
  if (!hist || !order) {
    if (owned) cudaFree(owned);
    return CAPY_ERR_OOM;
  }
  auto fail = [&](cudaError_t e, const char* what) {
    if (owned) cudaFree(owned);
    return cuda_fail(ctx, e, what);
  };
#define PLAN_CUDA(expr)                                  \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return fail(_e, #expr);       \
  } while (0)
  uint32_t* summary = hist + kLenBins;
  PLAN_CUDA(cudaMemsetAsync(hist, 0, (size_t)kLenBins * 4 + 64, st));
  len_hist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_off, n, stride_bytes, hist);
  len_scan_kernel<<<1, 1024, 0, st>>>(hist, summary);
  ctx->launches += 2;
  uint32_t h_sum[4] = {(uint32_t)h_total, (uint32_t)(h_total >> 32), h_max, h_bins};
  if (!h_off) {
    PLAN_CUDA(cudaMemcpyAsync(h_sum, summary, sizeof h_sum, cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaStreamSynchronize(st));
  }
  const uint64_t total_blocks = (uint64_t)h_sum[0] | ((uint64_t)h_sum[1] << 32);
  const uint32_t max_blocks = h_sum[2] + 1, bins = h_sum[3];
  // time in units of one thread-per-state permutation on one scheduler: all work spread over every scheduler
  // (32 states per warp, one warp per scheduler) vs the longest chain
  const double ideal = (double)total_blocks / (32.0 * 4.0 * dc.sm_count);
  // Chain-bound batch: the longest items go to the warp-per-item and the two-threads-per-item tiers
  if (allow_pair && max_blocks > 64 && (double)max_blocks > 1.05 * ideal) {
    std::vector<uint32_t> cum(kLenBins);  // after the scan hist[k] = number of items in bins > k
    if (h_off) {
      uint32_t above = 0;
      for (uint32_t k = kLenBins; k-- > 0;) {
        cum[k] = above;
        above += h_cnt[k];
      }
    } else {
      PLAN_CUDA(cudaMemcpyAsync(cum.data(), hist, (size_t)kLenBins * 4, cudaMemcpyDeviceToHost, st));
      PLAN_CUDA(cudaStreamSynchronize(st));
    }
    plan_tiers(cum, n, max_blocks, (double)total_blocks, dc.sm_count, plan->warp_items, &plan->pair_items, force_c);
  }
  if (bins == 1) plan->uniform_blocks = max_blocks;
  if (bins > 1) {  // (uniform lengths: nothing to order)
    len_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_off, n, stride_bytes, hist, order);
    ctx->launches++;
    PLAN_CUDA(cudaGetLastError());
    plan->order = order;
    plan->warps_per_smsp = 0;
    if ((double)max_blocks * 4.0 > 0.7 * ideal) {
      plan->warps_per_smsp = 3;
      if ((double)max_blocks * 3.0 > 0.7 * ideal) plan->warps_per_smsp = 2;
      if ((double)max_blocks * 2.0 > 0.7 * ideal) plan->warps_per_smsp = 1;
    }
  }
#undef PLAN_CUDA
  if (cached) {
    if (!plan->order) {  // uniform batch: the entry records "nothing to do"
      cudaFree(owned);
      owned = nullptr;
    }
    if (dc.plans.size() >= kMaxPlanEntries) {  // evict the least recently used plan
      size_t lru = 0;
      for (size_t k = 1; k < dc.plans.size(); k++)
        if (dc.plans[k].last_use < dc.plans[lru].last_use) lru = k;
      if (dc.plans[lru].owned_order) cudaFree(dc.plans[lru].owned_order);  // waits for work that still reads it
      dc.plans.erase(dc.plans.begin() + lru);
    }
    PlanEntry e;
    e.off = d_off;
    e.n = n;
    e.unit = stride_bytes;
    e.allow_pair = allow_pair;
    e.plan = *plan;
    e.owned_order = owned;
    e.last_use = ++dc.use_clock;
    dc.plans.push_back(e);
  }
  return CAPY_OK;
}

template <int LANES>
static int launch_sponge_t(capy_ctx* ctx, cudaStream_t stream, SpongeJob J, unsigned block, const LaunchPlan& plan) {
  const uint64_t warp_items = plan.warp_items[0] + plan.warp_items[1] + plan.warp_items[2], pair_items = plan.pair_items;
  if (pair_items || warp_items) {
    // chain-bound batch: warp blocks (one, two, three chains per scheduler), then pair blocks, then thread-per-item
    // blocks, one block per SM.  A block has 384 threads: a warp-tier block of class c uses 4 c warps, the other tiers
    // the first four; the rest exit at once.
    SpongeTiers tiers;
    uint32_t warp_blocks = 0;
    uint64_t rank = 0;
    for (int c = 0; c < 3; c++) {
      tiers.first_block[c] = warp_blocks;
      tiers.first_rank[c] = rank;
      warp_blocks += grid_for(plan.warp_items[c], 4 * (c + 1));
      rank += plan.warp_items[c];
    }
    tiers.warp_blocks = warp_blocks;
    tiers.pair_blocks = grid_for(2 * pair_items, 128);
    const unsigned solo_blocks = grid_for(J.n - pair_items - warp_items, 128);
    J.warp_items = warp_items;
    J.first = warp_items + pair_items;
    CAPY_CUDA(ctx, cudaFuncSetAttribute(sponge_tiered_kernel<LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    sponge_tiered_kernel<LANES><<<warp_blocks + tiers.pair_blocks + solo_blocks, 384, 226 * 1024, stream>>>(J, tiers);
    ctx->launches++;
    CAPY_CUDA(ctx, cudaGetLastError());
    return CAPY_OK;
  }
  size_t smem = 0;
  const int warps_per_smsp = plan.warps_per_smsp;
  if (warps_per_smsp >= 1 && warps_per_smsp <= 3) {
    // one 128-thread block = one warp per scheduler; W blocks per SM via the shared-memory footprint
    block = 128;
    smem = (size_t)(227 * 1024 / warps_per_smsp) - 1024;
    CAPY_CUDA(ctx, cudaFuncSetAttribute(sponge_kernel<LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  sponge_kernel<LANES><<<grid_for(J.n, block), block, smem, stream>>>(J);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}

// Block size for a one-thread-per-item kernel with few items: pick the block size whose block count
// spreads most evenly over the SMs (ties go to the larger block).
static unsigned pick_block(uint64_t n, int sm_count, int threads_per_sm) {
  unsigned best = 128;
  double best_eff = -1.0;
  for (unsigned bs : {128u, 64u, 32u}) {
    const uint64_t blocks = (n + bs - 1) / bs;
    const uint64_t slots = (uint64_t)sm_count * (threads_per_sm / bs);  // resident blocks per wave
    const uint64_t waves = (blocks + slots - 1) / slots;
    const double per_sm = (double)blocks / sm_count;
    const double busiest = waves > 1 ? (double)waves * (threads_per_sm / bs) : (double)((blocks + sm_count - 1) / sm_count);
    const double eff = per_sm / busiest;
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = bs;
    }
  }
  return best;
}

static int launch_sponge(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int lanes, const SpongeJob& J,
                         const LaunchPlan& plan = LaunchPlan()) {
  if (J.n == 0) return CAPY_OK;
  const unsigned block = pick_block(J.n, dc.sm_count, 384);
  switch (lanes) {
    case 9: return launch_sponge_t<9>(ctx, stream, J, block, plan);
    case 13: return launch_sponge_t<13>(ctx, stream, J, block, plan);
    case 17: return launch_sponge_t<17>(ctx, stream, J, block, plan);
    case 18: return launch_sponge_t<18>(ctx, stream, J, block, plan);
    case 19: return launch_sponge_t<19>(ctx, stream, J, block, plan);
    case 21: return launch_sponge_t<21>(ctx, stream, J, block, plan);
    default: return CAPY_ERR_BAD_ARG;
  }
}

// chain cutting of uniform batches (defined with the cSHAKE / KMAC launchers below)
static bool chain_cut(int sm_count, uint64_t n, uint64_t absorb_blocks, uint64_t squeeze_extra, uint64_t* cut);
static int launch_sponge_chain(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int lanes, const SpongeJob& J, uint64_t cut);

static SpongeJob empty_job() {
  SpongeJob J;
  memset(&J, 0, sizeof J);
  return J;
}

// ---- SHA3-d ------------------------------------------------------------------------------------
// tail_readable: the 8-byte word that holds the last message byte of the LAST row may be read in full (the host entry
// points stage into buffers with slack; a caller-owned device buffer may end with the last message byte)
static int launch_sha3(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int d, const uint8_t* data, const uint64_t* off,
                       uint64_t msg_len, uint64_t stride, uint64_t n, uint8_t* out, uint32_t flags = 0,
                       bool tail_readable = true) {
  if (!valid_secparam(d)) return CAPY_ERR_BAD_SECPARAM;
  if (n == 0) return CAPY_OK;
  if (!off && !tail_readable && (msg_len & 7u) != 0 && n > 1 && (reinterpret_cast<uintptr_t>(data) & 7u) == 0 &&
      (stride & 7u) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
    // every row but the last has its tail word inside its own stride or the next row; the last row goes through the
    // byte-granular kernel, which never reads past the message
    int rc = launch_sha3(ctx, dc, stream, d, data, nullptr, msg_len, stride, n - 1, out, flags, true);
    if (rc) return rc;
    return launch_sha3(ctx, dc, stream, d, data + (n - 1) * stride, nullptr, msg_len, stride, 1, out + (n - 1) * (uint64_t)(d / 8),
                       flags, false);
  }
  const uint32_t rate = (1600 - sha3_capacity(d)) / 8;  // 144 / 136 / 104 / 72
  const int lanes = (int)rate / 8;
  if (!off && (reinterpret_cast<uintptr_t>(data) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
      (stride & 15u) == 0 && (d == 256 || d == 512) && msg_len >= 16 && (msg_len & 15u) == 0 && msg_len < rate) {
    // compile-time block layout: single-block messages of a whole number of 16-byte rows (cfg 1 is 64 B at D256)
    const unsigned block = 128, grid = grid_for(n, block);
    const uint4* dq = reinterpret_cast<const uint4*>(data);
    uint4* oq = reinterpret_cast<uint4*>(out);
    const uint64_t s16 = stride / 16;
#define CAPY_SHORT(L, M, OB) sha3_short_kernel<L, M><<<grid, block, 0, stream>>>(dq, s16, 0x06u, oq, OB, n)
    if (d == 256) {
      switch (msg_len / 8) {
        case 2: CAPY_SHORT(17, 2, 32); break;
        case 4: CAPY_SHORT(17, 4, 32); break;
        case 6: CAPY_SHORT(17, 6, 32); break;
        case 8: CAPY_SHORT(17, 8, 32); break;
        case 10: CAPY_SHORT(17, 10, 32); break;
        case 12: CAPY_SHORT(17, 12, 32); break;
        case 14: CAPY_SHORT(17, 14, 32); break;
        default: CAPY_SHORT(17, 16, 32); break;
      }
    } else {
      switch (msg_len / 8) {
        case 2: CAPY_SHORT(9, 2, 64); break;
        case 4: CAPY_SHORT(9, 4, 64); break;
        case 6: CAPY_SHORT(9, 6, 64); break;
        default: CAPY_SHORT(9, 8, 64); break;
      }
    }
#undef CAPY_SHORT
    ctx->launches++;
    CAPY_CUDA(ctx, cudaGetLastError());
    return CAPY_OK;
  }
  uint64_t cut = 0;
  const bool cut_fixed = !off && chain_cut(dc.sm_count, n, msg_len / rate + 1, 0, &cut);  // long equal messages, uneven fill
  // the uniform kernel stores 4-byte granules (28- and 48-byte digests) and reads whole 8-byte words
  if (!off && !cut_fixed && (reinterpret_cast<uintptr_t>(data) & 7u) == 0 && (stride & 7u) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 3u) == 0 && (tail_readable || (msg_len & 7u) == 0)) {
    const uint32_t suffix = (msg_len % 136u == 135u) ? 0x86u : 0x06u;  // shake_functions.rs:25-29 (Q2)
    const unsigned block = 128, grid = grid_for(n, block);
    switch (lanes) {
      case 18: sha3_uniform_kernel<18><<<grid, block, 0, stream>>>(data, stride, msg_len, suffix, out, d / 8, n); break;
      case 17: sha3_uniform_kernel<17><<<grid, block, 0, stream>>>(data, stride, msg_len, suffix, out, d / 8, n); break;
      case 13: sha3_uniform_kernel<13><<<grid, block, 0, stream>>>(data, stride, msg_len, suffix, out, d / 8, n); break;
      case 9: sha3_uniform_kernel<9><<<grid, block, 0, stream>>>(data, stride, msg_len, suffix, out, d / 8, n); break;
    }
    ctx->launches++;
    CAPY_CUDA(ctx, cudaGetLastError());
    return CAPY_OK;
  }
  SpongeJob J = empty_job();
  J.data = data;
  J.off = off;
  J.msg_len = msg_len;
  J.msg_stride = stride;
  J.trailer_len = 1;
  J.sha3_suffix = 1;
  J.rate = rate;
  J.out = out;
  J.out_stride = J.out_bytes = (uint64_t)d / 8;
  J.sq_lanes = (1600 - d) / 64;  // Rate::from(&d), sponge.rs:27 (first d/8 bytes are all that is kept)
  J.n = n;
  LaunchPlan plan;
  if (off && !(flags & CAPY_FLAG_NO_SORT)) {
    int rc = plan_ragged(ctx, dc, stream, off, n, rate, &plan, !(flags & CAPY_FLAG_NO_PAIR),
                         (int)((flags >> CAPY_FLAG_WARP_COSCHED_SHIFT) & 3u));
    if (rc) return rc;
  }
  J.order = plan.order;
  if (cut_fixed) return launch_sponge_chain(ctx, dc, stream, lanes, J, cut);
  if (off && !plan.order && plan.uniform_blocks && chain_cut(dc.sm_count, n, plan.uniform_blocks, 0, &cut))
    return launch_sponge_chain(ctx, dc, stream, lanes, J, cut);
  return launch_sponge(ctx, dc, stream, lanes, J, plan);
}

// ---- cSHAKE / KMAC prefix -----------------------------------------------------------------------
static void host_left_encode(std::string& o, uint64_t v) {  // aux_functions.rs:34-49
  if (v == 0) {
    o.push_back(1);
    o.push_back(0);
    return;
  }
  int nb = 8;
  while (nb > 1 && ((v >> (8 * (nb - 1))) & 0xFF) == 0) nb--;
  o.push_back((char)nb);
  for (int i = nb - 1; i >= 0; i--) o.push_back((char)((v >> (8 * i)) & 0xFF));
}
static void host_encode_string(std::string& o, const uint8_t* s, size_t n) {  // aux_functions.rs:24-28
  host_left_encode(o, (uint64_t)n * 8);
  o.append(reinterpret_cast<const char*>(s), n);
}

static int get_prefix(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int d, const uint8_t* fn, uint32_t fn_len,
                      const uint8_t* cs, uint32_t cs_len, PrefixState** out) {
  std::string key;
  key.push_back((char)(d >> 8));
  key.push_back((char)d);
  host_left_encode(key, fn_len);
  key.append(reinterpret_cast<const char*>(fn), fn_len);
  key.append(reinterpret_cast<const char*>(cs), cs_len);
  auto it = dc.prefix_cache.find(key);
  if (it != dc.prefix_cache.end()) {
    it->second.last_use = ++dc.use_clock;
    *out = &it->second;
    return CAPY_OK;
  }
  // The customisation string is caller-controlled (compute_tagged_hash's S): bound the cache, least recently used out.
  // cudaFree waits for work that still reads the evicted buffers.
  if (dc.prefix_cache.size() >= kMaxPrefixEntries) {
    auto lru = dc.prefix_cache.begin();
    for (auto k = dc.prefix_cache.begin(); k != dc.prefix_cache.end(); ++k)
      if (k->second.last_use < lru->second.last_use) lru = k;
    if (lru->second.d_state) cudaFree(lru->second.d_state);
    if (lru->second.d_prefix) cudaFree(lru->second.d_prefix);
    dc.prefix_cache.erase(lru);
  }
  const uint32_t w = bytepad_value(d);
  std::string p;  // bytepad(encode_string(N) || encode_string(S), w)   shake_functions.rs:50-55
  host_left_encode(p, w);
  host_encode_string(p, fn, fn_len);
  host_encode_string(p, cs, cs_len);
  p.append(w - p.size() % w, '\0');  // aux_functions.rs:14-16 (quirk Q3)
  PrefixState ps;
  ps.prefix_len = (uint32_t)p.size();
  auto fail = [&](cudaError_t e, const char* what) {  // nothing is cached on failure: free what was allocated
    if (ps.d_prefix) cudaFree(ps.d_prefix);
    if (ps.d_state) cudaFree(ps.d_state);
    return cuda_fail(ctx, e, what);
  };
#define PREFIX_CUDA(expr)                          \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return fail(_e, #expr); \
  } while (0)
  PREFIX_CUDA(cudaMalloc(&ps.d_prefix, p.size()));
  PREFIX_CUDA(cudaMalloc(&ps.d_state, 25 * sizeof(uint64_t)));
  PREFIX_CUDA(cudaMemcpyAsync(ps.d_prefix, p.data(), p.size(), cudaMemcpyHostToDevice, stream));
  const int lanes = (int)(w * 8 / 64);  // 21 / 21 / 19 / 17
  if (w % 8 == 0) {
    // the prefix is a whole number of blocks: fold it into a cached state once
    ps.skip_blocks = ps.prefix_len / w;
    switch (lanes) {
      case 21: prefix_state_kernel<21><<<1, 32, 0, stream>>>(ps.d_prefix, ps.skip_blocks, ps.d_state); break;
      case 19: prefix_state_kernel<19><<<1, 32, 0, stream>>>(ps.d_prefix, ps.skip_blocks, ps.d_state); break;
      case 17: prefix_state_kernel<17><<<1, 32, 0, stream>>>(ps.d_prefix, ps.skip_blocks, ps.d_state); break;
    }
    ctx->launches++;
    PREFIX_CUDA(cudaGetLastError());
  } else {
    ps.skip_blocks = 0;  // D224: w = 172 is not lane aligned (quirk Q7), blocks straddle the prefix end
  }
  // the host string dies at return, and later calls may use the entry from other streams: wait for the build
  PREFIX_CUDA(cudaStreamSynchronize(stream));
#undef PREFIX_CUDA
  ps.last_use = ++dc.use_clock;
  auto ins = dc.prefix_cache.emplace(key, ps);
  *out = &ins.first->second;
  return CAPY_OK;
}

static void fill_cshake_common(SpongeJob& J, int d, const PrefixState* ps) {
  const uint32_t w = bytepad_value(d);
  J.prefix = ps->d_prefix;
  J.prefix_len = ps->prefix_len;
  J.init_state = ps->skip_blocks ? ps->d_state : nullptr;
  J.skip_blocks = ps->skip_blocks;
  J.w = w;
  J.rate = (1600 - d) / 8;  // capacity = d bits (shake_functions.rs:63, quirk Q7)
  J.sq_lanes = (1600 - d) / 64;
}

// ---- chain splitting of uniform batches (sponge.cuh: sponge_chain_kernel) --------------------------------------------
// Where to cut: the batch must fill the schedulers unevenly, be large enough that no job-1 block has to wait for its
// job-0 block (at least as many job-0 blocks as the GPU holds at once: CAPY_SPONGE_MINB per SM), and leave the two halves
// about equally long; the cut lies inside -- or at the end of -- the absorb phase.  absorb_blocks counts the blocks after
// skip_blocks, squeeze_extra the permutations between squeeze blocks.  (A wrong estimate costs balance, never
// correctness: an item that is shorter than the cut simply has nothing left to absorb in job 1.)
static bool chain_cut(int sm_count, uint64_t n, uint64_t absorb_blocks, uint64_t squeeze_extra, uint64_t* cut) {
  if (getenv("CAPY_NO_CHAIN_SPLIT")) return false;  // A/B switch for the probes
  if (sm_count < 1) return false;
  const uint64_t warps = (n + 31) / 32, sched = 4ull * (uint64_t)sm_count, blocks = (n + 127) / 128;
  if (blocks < (uint64_t)CAPY_SPONGE_MINB * (uint64_t)sm_count || warps > 8 * sched) return false;
  auto fill = [&](uint64_t w) {
    const double x = (double)w / (double)sched;
    return x / std::ceil(x);
  };
  if (fill(2 * warps) < fill(warps) + 0.04) return false;
  const uint64_t perms = absorb_blocks + squeeze_extra;
  if (perms < 12) return false;
  uint64_t k = (perms + 1) / 2;
  if (k > absorb_blocks) k = absorb_blocks;
  if (3 * k < perms) return false;  // the absorb phase is too short to carry a half
  *cut = k;
  return true;
}

template <int LANES>
static int launch_chain_t(capy_ctx* ctx, cudaStream_t stream, const SpongeChain& C, unsigned grid) {
  sponge_chain_kernel<LANES><<<grid, 128, 0, stream>>>(C);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}

static int chain_stream_index(const DeviceCtx& dc, cudaStream_t stream) {
  for (int k = 1; k < kNumStreams; k++)
    if (dc.streams[k] == stream) return k;
  return 0;
}

// two jobs over the same ranks in one launch, job 1 of an item after job 0 of that item (sponge_chain_kernel)
static int launch_chain_jobs(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int lanes, const SpongeJob& J0, const SpongeJob& J1,
                             bool data_dependent) {
  const uint32_t nb = (uint32_t)((J0.n + 127) / 128);
  uint32_t* sync = (uint32_t*)scratch_get(dc, 145 + 2 * chain_stream_index(dc, stream), ((size_t)nb + 1) * sizeof(uint32_t));
  if (!sync) return CAPY_ERR_OOM;
  CAPY_CUDA(ctx, cudaMemsetAsync(sync, 0, ((size_t)nb + 1) * sizeof(uint32_t), stream));
  SpongeChain C;
  C.j[0] = J0;
  C.j[1] = J1;
  C.sync = sync;
  C.blocks_per_job = nb;
  C.data_dependent = data_dependent ? 1u : 0u;
  switch (lanes) {
    case 9: return launch_chain_t<9>(ctx, stream, C, 2 * nb);
    case 13: return launch_chain_t<13>(ctx, stream, C, 2 * nb);
    case 17: return launch_chain_t<17>(ctx, stream, C, 2 * nb);
    case 18: return launch_chain_t<18>(ctx, stream, C, 2 * nb);
    case 19: return launch_chain_t<19>(ctx, stream, C, 2 * nb);
    case 21: return launch_chain_t<21>(ctx, stream, C, 2 * nb);
    default: return CAPY_ERR_BAD_ARG;
  }
}

// J (uniform batch, no order) as two dependent jobs cut after `cut` absorbed blocks
static int launch_sponge_chain(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int lanes, const SpongeJob& J, uint64_t cut) {
  uint64_t* states = (uint64_t*)scratch_get(dc, 144 + 2 * chain_stream_index(dc, stream), (size_t)J.n * 25 * sizeof(uint64_t));
  if (!states) return CAPY_ERR_OOM;
  SpongeJob J0 = J, J1 = J;
  J0.chain_out = states;
  J0.stop_block = J.skip_blocks + cut;
  J1.chain_in = states;
  J1.skip_blocks = (uint32_t)(J.skip_blocks + cut);
  return launch_chain_jobs(ctx, dc, stream, lanes, J0, J1, false);
}

// blocks a uniform cSHAKE / KMAC item absorbs after the cached prefix, estimated from the message blocks of the plan
static uint64_t absorb_blocks_estimate(const SpongeJob& J, uint64_t key_len, bool has_key, uint64_t msg_blocks_whole) {
  uint64_t blocks = msg_blocks_whole + 1;  // the message and the block that holds its tail, the trailer and the pad
  if (has_key) blocks += (key_len + 8) / J.w + 1;  // bytepad(encode_string(K), w)
  if (!J.skip_blocks) blocks += J.prefix_len / J.rate;
  return blocks;
}

static int launch_cshake(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, int d, const uint8_t* data,
                         const uint64_t* off, uint64_t n, const uint8_t* fn, uint32_t fn_len, const uint8_t* cs,
                         uint32_t cs_len, uint64_t out_bits, uint8_t* out) {
  if (!valid_secparam(d)) return CAPY_ERR_BAD_SECPARAM;
  if (n == 0 || out_bits / 8 == 0) return CAPY_OK;
  PrefixState* ps;
  int rc = get_prefix(ctx, dc, stream, d, fn, fn_len, cs, cs_len, &ps);
  if (rc) return rc;
  SpongeJob J = empty_job();
  fill_cshake_common(J, d, ps);
  J.data = data;
  J.off = off;
  J.trailer = 0x04;  // shake_functions.rs:57
  J.trailer_len = 1;
  if (fn_len == 0 && cs_len == 0) J.q4_rate = (1600 - sha3_capacity(d)) / 8;  // quirk Q4, :59-61
  J.out = out;
  J.out_stride = J.out_bytes = out_bits / 8;
  J.n = n;
  LaunchPlan plan;
  if (off) {
    rc = plan_ragged(ctx, dc, stream, off, n, J.rate & ~7u, &plan);
    if (rc) return rc;
  }
  J.order = plan.order;
  const int lanes = (int)(bytepad_value(d) * 8 / 64);
  if (off && !plan.order && plan.uniform_blocks && (J.rate & 7u) == 0) {  // uniform batch: see launch_kmac_xof
    const uint64_t absorb = absorb_blocks_estimate(J, 0, false, plan.uniform_blocks - 1);
    const uint64_t sq = 8ull * J.sq_lanes, squeeze_extra = (J.out_bytes + sq - 1) / sq - 1;
    uint64_t cut;
    if (chain_cut(dc.sm_count, n, absorb, squeeze_extra, &cut)) return launch_sponge_chain(ctx, dc, stream, lanes, J, cut);
  }
  return launch_sponge(ctx, dc, stream, lanes, J, plan);
}

static int build_kmac_job(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& a, SpongeJob* out) {
  static const uint8_t kKmac[4] = {'K', 'M', 'A', 'C'};
  PrefixState* ps;
  int rc = get_prefix(ctx, dc, stream, a.d_bits, kKmac, 4, a.custom, a.custom_len, &ps);
  if (rc) return rc;
  SpongeJob J = empty_job();
  fill_cshake_common(J, a.d_bits, ps);
  J.keys = a.keys;
  J.key_off = a.key_off;
  J.key_len = (uint32_t)a.key_len;
  J.key_stride = a.key_stride;
  J.data = a.data;
  J.off = a.off;
  J.msg_len = a.msg_len;
  J.msg_stride = a.msg_stride;
  J.trailer = 0x040100u;  // right_encode(0) = 00 01 (shake_functions.rs:86) then 04 (:57)
  J.trailer_len = 3;
  J.out = a.out;
  J.out_off = a.out_off;
  J.out_stride = a.out_stride;
  J.out_bytes = a.out_bytes;
  J.xor_in = a.xor_in;
  J.n = a.n;
  *out = J;
  return CAPY_OK;
}

static int plan_kmac(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& a, const SpongeJob& J, LaunchPlan* plan) {
  *plan = LaunchPlan();
  if (a.no_sort) return CAPY_OK;
  if (a.off) return plan_ragged(ctx, dc, stream, a.off, a.n, J.rate & ~7u, plan);
  // keystream shape (sha3/encryptable.rs:41): the work of an item is its squeeze length
  if (a.out_off) return plan_ragged(ctx, dc, stream, a.out_off, a.n, 8u * J.sq_lanes, plan);
  return CAPY_OK;
}

int launch_kmac_xof(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& a) {
  if (!valid_secparam(a.d_bits)) return CAPY_ERR_BAD_SECPARAM;
  if (a.n == 0) return CAPY_OK;
  if (!a.out_off && a.out_bytes == 0) return CAPY_OK;
  SpongeJob J;
  int rc = build_kmac_job(ctx, dc, stream, a, &J);
  if (rc) return rc;
  LaunchPlan plan;
  rc = plan_kmac(ctx, dc, stream, a, J, &plan);
  if (rc) return rc;
  J.order = plan.order;
  const int lanes = (int)(bytepad_value(a.d_bits) * 8 / 64);
  // uniform batch (fixed-length call, or a plan that found one block count): cut the chains when that fills the GPU better
  const uint64_t unit = J.rate & ~7u;
  const uint64_t msg_blocks = a.off ? plan.uniform_blocks : a.msg_len / unit + 1;
  if (!a.out_off && !a.xor_in && !plan.order && msg_blocks && (J.rate & 7u) == 0) {
    const uint64_t absorb = absorb_blocks_estimate(J, a.key_len, a.keys != nullptr, msg_blocks - 1);
    const uint64_t sq = 8ull * J.sq_lanes, squeeze_extra = a.out_bytes ? (a.out_bytes + sq - 1) / sq - 1 : 0;
    uint64_t cut;
    if (chain_cut(dc.sm_count, a.n, absorb, squeeze_extra, &cut)) return launch_sponge_chain(ctx, dc, stream, lanes, J, cut);
  }
  return launch_sponge(ctx, dc, stream, lanes, J, plan);
}

// Two independent KMACXOF passes over the same items (same d, same n) as ONE launch: warps alternate between the two
// jobs, so a batch whose single pass fills the schedulers unevenly (2^16 items = 3.46 warps per scheduler, the last
// wave 15 % full) runs as 6.92 warps per scheduler.  The authenticated-encryption seal is such a pair: the tag pass
// absorbs the message, the keystream pass squeezes into the XOR (sha3/encryptable.rs:36-42).  Both jobs follow the
// plan of `a`.  Chain-bound batches (tiers) run the two passes one after the other.
template <int LANES>
static int launch_sponge2_t(capy_ctx* ctx, cudaStream_t stream, const SpongeJob& A, const SpongeJob& B, unsigned block,
                            const LaunchPlan& plan) {
  SpongeJob2 JJ;
  JJ.j[0] = A;
  JJ.j[1] = B;
  size_t smem = 0;
  if (plan.warps_per_smsp >= 1 && plan.warps_per_smsp <= 3) {
    block = 128;
    smem = (size_t)(227 * 1024 / plan.warps_per_smsp) - 1024;
    CAPY_CUDA(ctx, cudaFuncSetAttribute(sponge_kernel2<LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  const uint64_t warps = 2 * ((A.n + 31) / 32);
  sponge_kernel2<LANES><<<grid_for(warps * 32, block), block, smem, stream>>>(JJ);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}

int launch_kmac_xof2(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& a, const KmacDevArgs& b) {
  if (!valid_secparam(a.d_bits) || a.d_bits != b.d_bits || a.n != b.n) return CAPY_ERR_BAD_ARG;
  if (a.n == 0) return CAPY_OK;
  SpongeJob JA, JB;
  int rc = build_kmac_job(ctx, dc, stream, a, &JA);
  if (rc) return rc;
  rc = build_kmac_job(ctx, dc, stream, b, &JB);
  if (rc) return rc;
  LaunchPlan plan;
  rc = plan_kmac(ctx, dc, stream, a, JA, &plan);
  if (rc) return rc;
  JA.order = JB.order = plan.order;
  const int lanes = (int)(bytepad_value(a.d_bits) * 8 / 64);
  if (plan.warp_items[0] || plan.warp_items[1] || plan.warp_items[2] || plan.pair_items) {
    rc = launch_sponge(ctx, dc, stream, lanes, JA, plan);
    if (rc) return rc;
    return launch_sponge(ctx, dc, stream, lanes, JB, plan);
  }
  const unsigned block = pick_block(2 * a.n, dc.sm_count, 384);
  switch (lanes) {
    case 17: return launch_sponge2_t<17>(ctx, stream, JA, JB, block, plan);
    case 19: return launch_sponge2_t<19>(ctx, stream, JA, JB, block, plan);
    case 21: return launch_sponge2_t<21>(ctx, stream, JA, JB, block, plan);
    default: return CAPY_ERR_BAD_ARG;
  }
}

// Two KMACXOF passes where the SECOND absorbs what the FIRST wrote for the same item (authenticated decryption: the
// keystream pass recovers the plaintext, the tag pass runs over it -- sha3/encryptable.rs:66-75): one launch of
// dependent jobs instead of two kernels, so the partial wave of the first pass is filled by the second.  Both follow the
// plan (work order) of `second`.  Small or chain-bound batches run the two passes one after the other.
int launch_kmac_xof_dep(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& first, const KmacDevArgs& second) {
  if (!valid_secparam(first.d_bits) || first.d_bits != second.d_bits || first.n != second.n) return CAPY_ERR_BAD_ARG;
  if (first.n == 0) return CAPY_OK;
  SpongeJob J0, J1;
  int rc = build_kmac_job(ctx, dc, stream, first, &J0);
  if (rc) return rc;
  rc = build_kmac_job(ctx, dc, stream, second, &J1);
  if (rc) return rc;
  LaunchPlan plan;
  rc = plan_kmac(ctx, dc, stream, second, J1, &plan);
  if (rc) return rc;
  J0.order = J1.order = plan.order;
  const int lanes = (int)(bytepad_value(first.d_bits) * 8 / 64);
  const bool tiers = plan.warp_items[0] || plan.warp_items[1] || plan.warp_items[2] || plan.pair_items;
  const uint64_t blocks = (first.n + 127) / 128;
  if (tiers || (J0.rate & 7u) != 0 || blocks < (uint64_t)CAPY_SPONGE_MINB * (uint64_t)dc.sm_count || getenv("CAPY_NO_CHAIN_SPLIT")) {
    rc = launch_sponge(ctx, dc, stream, lanes, J0, plan);  // (a job-1 block of a small batch would wait for its job-0 block)
    if (rc) return rc;
    return launch_sponge(ctx, dc, stream, lanes, J1, plan);
  }
  return launch_chain_jobs(ctx, dc, stream, lanes, J0, J1, true);
}

}  // namespace capy

using namespace capy;

extern "C" {

// =================================================================================================
// diagnostics
// =================================================================================================
int capy_plan_tiers(const uint32_t* items_longer_than, uint32_t n_bins, uint64_t n, uint32_t max_blocks, uint64_t total_blocks,
                    int sm_count, uint64_t* warp_items, uint64_t* pair_items) {
  uint64_t w[3];
  int rc = capy_plan_tiers3(items_longer_than, n_bins, n, max_blocks, total_blocks, sm_count, w, pair_items);
  if (rc == CAPY_OK && warp_items) *warp_items = w[0] + w[1] + w[2];
  return warp_items ? rc : CAPY_ERR_BAD_ARG;
}

int capy_plan_tiers3(const uint32_t* items_longer_than, uint32_t n_bins, uint64_t n, uint32_t max_blocks, uint64_t total_blocks,
                     int sm_count, uint64_t* warp_items_by_sharing, uint64_t* pair_items) {
  if (!items_longer_than || !n_bins || !warp_items_by_sharing || !pair_items || sm_count < 2) return CAPY_ERR_BAD_ARG;
  std::vector<uint32_t> cum(items_longer_than, items_longer_than + n_bins);
  plan_tiers(cum, n, max_blocks, (double)total_blocks, sm_count, warp_items_by_sharing, pair_items);
  return CAPY_OK;
}

int capy_chain_cut(int sm_count, uint64_t n, uint64_t absorb_blocks, uint64_t squeeze_extra, uint64_t* cut_after_blocks) {
  if (!cut_after_blocks) return CAPY_ERR_BAD_ARG;
  *cut_after_blocks = 0;
  return chain_cut(sm_count, n, absorb_blocks, squeeze_extra, cut_after_blocks) ? 1 : 0;
}

int capy_lpt_shares(const uint64_t* off, uint64_t n, uint32_t parts, uint32_t unit_bytes, uint64_t per_item_cost,
                    uint32_t* owner) {
  if (!off || !owner || parts == 0 || unit_bytes == 0) return CAPY_ERR_BAD_ARG;
  auto shares = lpt_shares(off, n, parts, unit_bytes, per_item_cost);
  for (size_t d = 0; d < shares.size(); d++)
    for (const Run& r : shares[d].runs)
      for (uint64_t i = r.i0; i < r.i1; i++) owner[i] = (uint32_t)d;
  return CAPY_OK;
}

// =================================================================================================
// SHA3-d
// =================================================================================================
int capy_sha3_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_data,
                        const uint64_t* d_off, uint64_t n, uint8_t* d_digests, uint32_t flags) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size() || (n && (!d_data || !d_off || !d_digests)))
    return CAPY_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(*ctx->devs[dev_index].mu);  // host-side state of this device (scratch slots, caches)
  DeviceGuard g(ctx->devs[dev_index].dev);
  return launch_sha3(ctx, ctx->devs[dev_index], (cudaStream_t)stream, d_bits, d_data, d_off, 0, 0, n, d_digests, flags);
}

int capy_sha3_batch_fixed_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_data,
                              uint64_t msg_len, uint64_t stride, uint64_t n, uint8_t* d_digests, uint32_t) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size() || (n && (!d_data || !d_digests)) || stride < msg_len)
    return CAPY_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(*ctx->devs[dev_index].mu);  // host-side state of this device (scratch slots, caches)
  DeviceGuard g(ctx->devs[dev_index].dev);
  return launch_sha3(ctx, ctx->devs[dev_index], (cudaStream_t)stream, d_bits, d_data, nullptr, msg_len, stride, n, d_digests, 0,
                     /*tail_readable=*/false);
}

int capy_sha3_batch(capy_ctx* ctx, int d_bits, const uint8_t* data, const uint64_t* off, uint64_t n, uint8_t* digests,
                    uint32_t flags) {
  if (!ctx || (n && (!data || !off || !digests))) return CAPY_ERR_BAD_ARG;
  if (!valid_secparam(d_bits)) return CAPY_ERR_BAD_SECPARAM;
  if (n == 0) return CAPY_OK;
  const size_t ob = (size_t)d_bits / 8;
  if (ctx->devs.size() > 1) {
    // several devices: the outliers of the batch (long chains) are dealt out longest first, the rest follows in
    // contiguous ranges (lpt_shares).  When nothing had to be dealt out every share is one run and the plain path
    // below does the same thing with pipelined chunks.
    const uint32_t rate = (1600 - sha3_capacity(d_bits)) / 8;
    auto shares = lpt_shares(off, n, ctx->devs.size(), rate, 2);
    bool scattered = false;
    for (const DeviceShare& s : shares) scattered |= s.runs.size() > 1;
    if (scattered) {
      std::vector<Range> slots;
      for (size_t k = 0; k < shares.size(); k++) slots.push_back({k, k + 1});  // for_each_device: one closure per device
      return for_each_device(ctx, slots, [&](DeviceCtx& dc, Range slot) -> int {
        const DeviceShare& sh = shares[slot.i0];
        if (sh.items == 0) return CAPY_OK;
        cudaStream_t st = dc.streams[0];
        uint8_t* d_data = (uint8_t*)scratch_get(dc, 0, (size_t)sh.bytes + 16);
        uint64_t* d_off = (uint64_t*)scratch_get(dc, 1, (size_t)(sh.items + 1) * 8);
        uint8_t* d_out = (uint8_t*)scratch_get(dc, 2, (size_t)sh.items * ob);
        if (!d_data || !d_off || !d_out) return CAPY_ERR_OOM;
        std::vector<uint64_t> loc(sh.items + 1);  // offsets of this device's items in its own packed buffer
        uint64_t at = 0, j = 0;
        for (const Run& r : sh.runs) {  // one copy per run of consecutive messages
          const uint64_t b0 = off[r.i0], b1 = off[r.i1];
          if (b1 > b0) {
            const int rc = copy_in(ctx, dc, st, 0, d_data + at, data + b0, (size_t)(b1 - b0));
            if (rc) return rc;
          }
          for (uint64_t i = r.i0; i < r.i1; i++) loc[j++] = at + (off[i] - b0);
          at += b1 - b0;
        }
        loc[j] = at;
        CAPY_CUDA(ctx, cudaMemcpyAsync(d_off, loc.data(), (size_t)(sh.items + 1) * 8, cudaMemcpyHostToDevice, st));
        host_off_register(dc, d_off, loc.data(), sh.items + 1);
        int rc = launch_sha3(ctx, dc, st, d_bits, d_data, d_off, 0, 0, sh.items, d_out, flags);
        if (rc) return rc;
        j = 0;
        for (const Run& r : sh.runs) {  // the digests of a run are consecutive on both sides
          rc = copy_out(ctx, dc, st, 2, digests + r.i0 * ob, d_out + j * ob, (size_t)(r.i1 - r.i0) * ob);
          if (rc) return rc;
          j += r.i1 - r.i0;
        }
        CAPY_CUDA(ctx, cudaStreamSynchronize(st));  // (`loc` stays alive until here)
        return CAPY_OK;
      });
    }
  }
  auto shards = split_items(off, 0, 0, n, ctx->devs.size(), 200);
  return for_each_device(ctx, shards, [&](DeviceCtx& dc, Range sh) -> int {
    // long messages: few big chunks, so that the longest-first schedule sees enough work beside them
    const size_t nchunks = ragged_chunk_count(off, sh.i0, sh.i1, 0);
    auto chunks = split_items(off, 0, sh.i0, sh.i1, nchunks, 200);
    for (size_t c = 0; c < chunks.size(); c++) {
      const int s = (int)(c % kNumStreams);
      cudaStream_t st = dc.streams[s];
      const Range ch = chunks[c];
      StagedPacked sp;
      int rc = stage_packed(ctx, dc, st, 3 * s, 3 * s + 1, data, off, ch.i0, ch.i1, &sp);
      if (rc) return rc;
      uint8_t* d_out = (uint8_t*)scratch_get(dc, 3 * s + 2, (size_t)(ch.i1 - ch.i0) * ob);
      if (!d_out) return CAPY_ERR_OOM;
      rc = launch_sha3(ctx, dc, st, d_bits, sp.d_base, sp.d_off, 0, 0, ch.i1 - ch.i0, d_out, flags);
      if (rc) return rc;
      rc = copy_out(ctx, dc, st, 3 * s + 2, digests + ch.i0 * ob, d_out, (size_t)(ch.i1 - ch.i0) * ob);
      if (rc) return rc;
    }
    for (int s = 0; s < kNumStreams; s++) CAPY_CUDA(ctx, cudaStreamSynchronize(dc.streams[s]));
    return CAPY_OK;
  });
}

int capy_sha3_batch_fixed(capy_ctx* ctx, int d_bits, const uint8_t* data, uint64_t msg_len, uint64_t stride, uint64_t n,
                          uint8_t* digests, uint32_t) {
  if (!ctx || (n && (!data || !digests)) || stride < msg_len) return CAPY_ERR_BAD_ARG;
  if (!valid_secparam(d_bits)) return CAPY_ERR_BAD_SECPARAM;
  if (n == 0) return CAPY_OK;
  const size_t ob = (size_t)d_bits / 8;
  // the uniform kernel wants 8-byte aligned rows; the copy lands on a 256-byte aligned buffer, so
  // only the stride matters
  auto shards = split_items(nullptr, stride, 0, n, ctx->devs.size(), 0);
  return for_each_device(ctx, shards, [&](DeviceCtx& dc, Range sh) -> int {
    auto chunks = split_items(nullptr, stride, sh.i0, sh.i1, chunk_count((sh.i1 - sh.i0) * stride, sh.i1 - sh.i0), 0);
    for (size_t c = 0; c < chunks.size(); c++) {
      const int s = (int)(c % kNumStreams);
      cudaStream_t st = dc.streams[s];
      const Range ch = chunks[c];
      const uint64_t cnt = ch.i1 - ch.i0;
      // last row may be shorter than the stride in the caller's buffer
      const size_t in_bytes = (size_t)((cnt - 1) * stride + msg_len);
      uint8_t* d_in = (uint8_t*)scratch_get(dc, 3 * s, in_bytes + 16);
      uint8_t* d_out = (uint8_t*)scratch_get(dc, 3 * s + 2, (size_t)cnt * ob);
      if (!d_in || !d_out) return CAPY_ERR_OOM;
      int rc = copy_in(ctx, dc, st, 3 * s, d_in, data + ch.i0 * stride, in_bytes);
      if (rc) return rc;
      rc = launch_sha3(ctx, dc, st, d_bits, d_in, nullptr, msg_len, stride, cnt, d_out);
      if (rc) return rc;
      rc = copy_out(ctx, dc, st, 3 * s + 2, digests + ch.i0 * ob, d_out, (size_t)cnt * ob);
      if (rc) return rc;
    }
    for (int s = 0; s < kNumStreams; s++) CAPY_CUDA(ctx, cudaStreamSynchronize(dc.streams[s]));
    return CAPY_OK;
  });
}

// =================================================================================================
// cSHAKE
// =================================================================================================
int capy_cshake_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_data,
                          const uint64_t* d_off, uint64_t n, const uint8_t* fn_name, uint32_t fn_len,
                          const uint8_t* custom, uint32_t custom_len, uint64_t out_bits, uint8_t* d_out) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size() || (n && (!d_data || !d_off || !d_out)) ||
      (fn_len && !fn_name) || (custom_len && !custom))
    return CAPY_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(*ctx->devs[dev_index].mu);  // host-side state of this device (scratch slots, caches)
  DeviceCtx& dc = ctx->devs[dev_index];
  DeviceGuard g(dc.dev);
  return launch_cshake(ctx, dc, (cudaStream_t)stream, d_bits, d_data, d_off, n, fn_name, fn_len, custom, custom_len,
                       out_bits, d_out);
}

int capy_cshake_batch(capy_ctx* ctx, int d_bits, const uint8_t* data, const uint64_t* off, uint64_t n,
                      const uint8_t* fn_name, uint32_t fn_len, const uint8_t* custom, uint32_t custom_len,
                      uint64_t out_bits, uint8_t* out) {
  if (!ctx || (n && (!data || !off || !out)) || (fn_len && !fn_name) || (custom_len && !custom)) return CAPY_ERR_BAD_ARG;
  if (!valid_secparam(d_bits)) return CAPY_ERR_BAD_SECPARAM;
  const size_t ob = (size_t)(out_bits / 8);
  if (n == 0 || ob == 0) return CAPY_OK;
  auto shards = split_items(off, 0, 0, n, ctx->devs.size(), 200 + ob);
  return for_each_device(ctx, shards, [&](DeviceCtx& dc, Range sh) -> int {
    auto chunks = split_items(off, 0, sh.i0, sh.i1,
                              ragged_chunk_count(off, sh.i0, sh.i1, (sh.i1 - sh.i0) * ob), 200 + ob);
    for (size_t c = 0; c < chunks.size(); c++) {
      const int s = (int)(c % kNumStreams);
      cudaStream_t st = dc.streams[s];
      const Range ch = chunks[c];
      StagedPacked sp;
      int rc = stage_packed(ctx, dc, st, 3 * s, 3 * s + 1, data, off, ch.i0, ch.i1, &sp);
      if (rc) return rc;
      uint8_t* d_out = (uint8_t*)scratch_get(dc, 3 * s + 2, (size_t)(ch.i1 - ch.i0) * ob);
      if (!d_out) return CAPY_ERR_OOM;
      rc = launch_cshake(ctx, dc, st, d_bits, sp.d_base, sp.d_off, ch.i1 - ch.i0, fn_name, fn_len, custom, custom_len,
                         out_bits, d_out);
      if (rc) return rc;
      rc = copy_out(ctx, dc, st, 3 * s + 2, out + ch.i0 * ob, d_out, (size_t)(ch.i1 - ch.i0) * ob);
      if (rc) return rc;
    }
    for (int s = 0; s < kNumStreams; s++) CAPY_CUDA(ctx, cudaStreamSynchronize(dc.streams[s]));
    return CAPY_OK;
  });
}

// =================================================================================================
// KMACXOF
// =================================================================================================
int capy_kmac_xof_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_keys,
                            const uint64_t* d_key_off, const uint8_t* d_data, const uint64_t* d_off, uint64_t n,
                            const uint8_t* custom, uint32_t custom_len, uint64_t out_bits, const uint64_t* d_out_off,
                            uint8_t* d_out) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size() ||
      (n && (!d_keys || !d_key_off || !d_data || !d_off || !d_out)) || (custom_len && !custom))
    return CAPY_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(*ctx->devs[dev_index].mu);  // host-side state of this device (scratch slots, caches)
  DeviceCtx& dc = ctx->devs[dev_index];
  DeviceGuard g(dc.dev);
  KmacDevArgs a{};
  a.d_bits = d_bits;
  a.keys = d_keys;
  a.key_off = d_key_off;
  a.data = d_data;
  a.off = d_off;
  a.n = n;
  a.custom = custom;
  a.custom_len = custom_len;
  a.out_bytes = a.out_stride = out_bits / 8;
  a.out_off = d_out_off;
  a.out = d_out;
  return launch_kmac_xof(ctx, dc, (cudaStream_t)stream, a);
}

int capy_kmac_xof_batch_fixed_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_keys,
                                  uint64_t key_len, uint64_t key_stride, const uint8_t* d_data, uint64_t msg_len,
                                  uint64_t msg_stride, uint64_t n, const uint8_t* custom, uint32_t custom_len,
                                  uint64_t out_bits, uint8_t* d_out) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size() || (n && (!d_keys || !d_out)) ||
      (n && msg_len && !d_data) || (custom_len && !custom) || key_stride < key_len || msg_stride < msg_len)
    return CAPY_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(*ctx->devs[dev_index].mu);  // host-side state of this device (scratch slots, caches)
  DeviceCtx& dc = ctx->devs[dev_index];
  DeviceGuard g(dc.dev);
  KmacDevArgs a{};
  a.d_bits = d_bits;
  a.keys = d_keys;
  a.key_len = key_len;
  a.key_stride = key_stride;
  a.data = d_data ? d_data : d_keys;  // never dereferenced when msg_len == 0
  a.msg_len = msg_len;
  a.msg_stride = msg_stride;
  a.n = n;
  a.custom = custom;
  a.custom_len = custom_len;
  a.out_bytes = a.out_stride = out_bits / 8;
  a.out = d_out;
  return launch_kmac_xof(ctx, dc, (cudaStream_t)stream, a);
}

int capy_kmac_xof_batch(capy_ctx* ctx, int d_bits, const uint8_t* keys, const uint64_t* key_off, const uint8_t* data,
                        const uint64_t* off, uint64_t n, const uint8_t* custom, uint32_t custom_len, uint64_t out_bits,
                        const uint64_t* out_off, uint8_t* out) {
  if (!ctx || (n && (!keys || !key_off || !data || !off || !out)) || (custom_len && !custom)) return CAPY_ERR_BAD_ARG;
  if (!valid_secparam(d_bits)) return CAPY_ERR_BAD_SECPARAM;
  const size_t ob = (size_t)(out_bits / 8);
  if (n == 0 || (!out_off && ob == 0)) return CAPY_OK;
  auto out_begin = [&](uint64_t i) -> uint64_t { return out_off ? out_off[i] : i * ob; };
  auto shards = split_items(off, 0, 0, n, ctx->devs.size(), 400 + ob);
  return for_each_device(ctx, shards, [&](DeviceCtx& dc, Range sh) -> int {
    const uint64_t tot = off[sh.i1] - off[sh.i0] + (out_begin(sh.i1) - out_begin(sh.i0));
    auto chunks = split_items(off, 0, sh.i0, sh.i1, ragged_chunk_count(off, sh.i0, sh.i1, tot - (off[sh.i1] - off[sh.i0])), 400 + ob);
    // 5 scratch slots per stream: data, off, keys, key_off, out (+ out_off)
    for (size_t c = 0; c < chunks.size(); c++) {
      const int s = (int)(c % kNumStreams);
      cudaStream_t st = dc.streams[s];
      const Range ch = chunks[c];
      const uint64_t cnt = ch.i1 - ch.i0;
      StagedPacked sx, sk;
      int rc = stage_packed(ctx, dc, st, 6 * s, 6 * s + 1, data, off, ch.i0, ch.i1, &sx);
      if (rc) return rc;
      rc = stage_packed(ctx, dc, st, 6 * s + 2, 6 * s + 3, keys, key_off, ch.i0, ch.i1, &sk);
      if (rc) return rc;
      const uint64_t ob0 = out_begin(ch.i0), ob1 = out_begin(ch.i1);
      uint8_t* d_out = (uint8_t*)scratch_get(dc, 6 * s + 4, (size_t)(ob1 - ob0) + 16);
      if (!d_out) return CAPY_ERR_OOM;
      const uint64_t* d_out_off = nullptr;
      uint8_t* d_out_base = d_out;
      if (out_off) {
        uint64_t* p = (uint64_t*)scratch_get(dc, 6 * s + 5, (size_t)(cnt + 1) * 8);
        if (!p) return CAPY_ERR_OOM;
        CAPY_CUDA(ctx, cudaMemcpyAsync(p, out_off + ch.i0, (size_t)(cnt + 1) * 8, cudaMemcpyHostToDevice, st));
        d_out_off = p;
        d_out_base = reinterpret_cast<uint8_t*>(reinterpret_cast<uintptr_t>(d_out) - (uintptr_t)ob0);
      }
      KmacDevArgs a{};
      a.d_bits = d_bits;
      a.keys = sk.d_base;
      a.key_off = sk.d_off;
      a.data = sx.d_base;
      a.off = sx.d_off;
      a.n = cnt;
      a.custom = custom;
      a.custom_len = custom_len;
      a.out_bytes = a.out_stride = ob;
      a.out_off = d_out_off;
      a.out = d_out_base;
      rc = launch_kmac_xof(ctx, dc, st, a);
      if (rc) return rc;
      if (ob1 > ob0) {
        rc = copy_out(ctx, dc, st, 6 * s + 4, out + ob0, d_out, (size_t)(ob1 - ob0));
        if (rc) return rc;
      }
    }
    for (int s = 0; s < kNumStreams; s++) CAPY_CUDA(ctx, cudaStreamSynchronize(dc.streams[s]));
    return CAPY_OK;
  });
}

// =================================================================================================
// FIPS 202 SHAKE (no reference counterpart)
// =================================================================================================
int capy_fips_shake_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int shake_bits, const uint8_t* d_data,
                              const uint64_t* d_off, uint64_t n, uint64_t out_bytes, uint8_t* d_out) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size() || (n && (!d_data || !d_off || !d_out)) ||
      (shake_bits != 128 && shake_bits != 256))
    return CAPY_ERR_BAD_ARG;
  if (n == 0 || out_bytes == 0) return CAPY_OK;
  std::lock_guard<std::mutex> lk(*ctx->devs[dev_index].mu);  // host-side state of this device (scratch slots, caches)
  DeviceGuard g(ctx->devs[dev_index].dev);
  SpongeJob J = empty_job();
  J.data = d_data;
  J.off = d_off;
  J.trailer = 0x1F;
  J.trailer_len = 1;
  J.fips_pad = 1;
  J.rate = (1600 - 2 * shake_bits) / 8;
  J.sq_lanes = J.rate / 8;
  J.out = d_out;
  J.out_stride = J.out_bytes = out_bytes;
  J.n = n;
  DeviceCtx& dc = ctx->devs[dev_index];
  LaunchPlan plan;  // longest first, tiers for chain-bound batches, chains of a uniform batch cut in two: as for cSHAKE
  int rc = plan_ragged(ctx, dc, (cudaStream_t)stream, d_off, n, J.rate, &plan);
  if (rc) return rc;
  J.order = plan.order;
  const int lanes = (int)J.rate / 8;
  if (!plan.order && plan.uniform_blocks) {
    const uint64_t sq = 8ull * J.sq_lanes;
    uint64_t cut;
    if (chain_cut(dc.sm_count, n, plan.uniform_blocks, (out_bytes + sq - 1) / sq - 1, &cut))
      return launch_sponge_chain(ctx, dc, (cudaStream_t)stream, lanes, J, cut);
  }
  return launch_sponge(ctx, dc, (cudaStream_t)stream, lanes, J, plan);
}

}  // extern "C"
