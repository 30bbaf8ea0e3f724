// keccak_sweep.cu -- tuning tool (not part of the library): SHA3-256 over N x 64-byte messages
// with the round loop unrolled U times and B threads per block; prints perms/s per variant.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "keccak.cuh"
using namespace capy;

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e = (x);                                                                      \
    if (e != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                                \
    }                                                                                         \
  } while (0)

template <int U, int B, int MINB, int PERMS, int ACTIVE = 32>
__global__ void __launch_bounds__(B, MINB) k(const uint4* __restrict__ in, uint4* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if ((threadIdx.x & 31) >= ACTIVE) return;  // half-populated warps: does a 16-lane warp issue faster?
  Lane a[25];
  state_zero(a);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint4 v = __ldg(in + 4 * (size_t)i + j);
    a[2 * j].lo = v.x; a[2 * j].hi = v.y; a[2 * j + 1].lo = v.z; a[2 * j + 1].hi = v.w;
  }
  a[8].lo ^= 0x06u;
  a[16].hi ^= 0x80000000u;
#pragma unroll 1
  for (int p = 0; p < PERMS; p++) keccak_f1600<U>(a);
  out[2 * (size_t)i] = make_uint4(a[0].lo, a[0].hi, a[1].lo, a[1].hi);
  out[2 * (size_t)i + 1] = make_uint4(a[2].lo, a[2].hi, a[3].lo, a[3].hi);
}

template <int U, int B, int MINB, int PERMS, int ACTIVE = 32>
static void run(const uint4* in, uint4* out, int n, int blocks_per_sm = 0) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int grid = (n + B - 1) / B;
  size_t smem = 0;
  if (blocks_per_sm > 0) {
    smem = (size_t)(227 * 1024 / blocks_per_sm) - 1024;
    CK(cudaFuncSetAttribute(k<U, B, MINB, PERMS, ACTIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  k<U, B, MINB, PERMS, ACTIVE><<<grid, B, smem>>>(in, out, n);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0));
    k<U, B, MINB, PERMS, ACTIVE><<<grid, B, smem>>>(in, out, n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k<U, B, MINB, PERMS, ACTIVE>));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<U, B, MINB, PERMS, ACTIVE>, B, smem));
  double perms = (double)n * PERMS / (best * 1e-3);
  printf("{\"unroll\": %d, \"block\": %d, \"minb\": %d, \"perms_per_thread\": %d, \"regs\": %d, \"warps_per_sm\": %d, "
         "\"ms\": %.4f, \"gperm_s\": %.3f, \"tops_4320\": %.3f, \"active_lanes\": %d}\n",
         U, B, MINB, PERMS, fa.numRegs, occ * B / 32, best, perms / 1e9, perms * 4320 / 1e12, ACTIVE);
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : (1 << 20);
  uint4 *in, *out;
  CK(cudaMalloc(&in, (size_t)n * 64));
  CK(cudaMalloc(&out, (size_t)n * 32));
  std::vector<unsigned> h((size_t)n * 16);
  unsigned x = 12345;
  for (auto& v : h) { x = x * 1664525u + 1013904223u; v = x; }
  CK(cudaMemcpy(in, h.data(), (size_t)n * 64, cudaMemcpyHostToDevice));
  // 1 permutation per thread (cfg 1 shape)
  run<1, 128, 1, 1>(in, out, n);
  run<2, 128, 1, 1>(in, out, n);
  run<4, 128, 1, 1>(in, out, n);
  run<24, 128, 1, 1>(in, out, n);
  run<1, 64, 1, 1>(in, out, n);
  run<1, 256, 1, 1>(in, out, n);
  run<2, 256, 1, 1>(in, out, n);
  run<1, 128, 7, 1>(in, out, n);
  run<2, 128, 7, 1>(in, out, n);
  // 32 permutations per thread (long-message shape: permutation throughput without load/store)
  run<1, 128, 1, 32>(in, out, n);
  run<2, 128, 1, 32>(in, out, n);
  run<4, 128, 1, 32>(in, out, n);
  run<24, 128, 1, 32>(in, out, n);
  run<1, 256, 1, 32>(in, out, n);
  run<2, 64, 1, 32>(in, out, n);
  // occupancy-throttled runs (one / two / three 128-thread blocks per SM = 1 / 2 / 3 warps per scheduler):
  // how fast does a single sponge chain advance when it has the scheduler (almost) to itself?
  run<1, 128, 1, 32>(in, out, n, 1);
  run<2, 128, 1, 32>(in, out, n, 1);
  run<4, 128, 1, 32>(in, out, n, 1);
  run<24, 128, 1, 32>(in, out, n, 1);
  run<1, 128, 1, 32>(in, out, n, 2);
  run<4, 128, 1, 32>(in, out, n, 2);
  run<24, 128, 1, 32>(in, out, n, 2);
  run<1, 128, 1, 32>(in, out, n, 3);
  run<24, 128, 1, 32>(in, out, n, 3);
  run<1, 128, 1, 32, 16>(in, out, n, 1);
  run<1, 128, 1, 32, 8>(in, out, n, 1);
  run<1, 128, 1, 32, 16>(in, out, n, 0);
  return 0;
}
