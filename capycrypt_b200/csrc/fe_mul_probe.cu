// fe_mul_probe.cu -- measurement tool: chains of field multiplications / squarings, integer (IMAD.WIDE) vs FP64 + IMAD, at W warps per scheduler
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fp448_f64.cuh"
using namespace capy;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

template <int KIND>
__global__ void __launch_bounds__(128) k_chain(const uint32_t* in, uint32_t* out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fe a, r;
#pragma unroll
  for (int i = 0; i < 16; i++) { a.v[i] = in[i * 128 + (t & 127)] & M28; r.v[i] = in[(16 + i) * 128 + (t & 127)] & M28; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (KIND == 0) fe_mul_inl(r, r, a);
    if (KIND == 1) fe_mul_f64(r, r, a);
    if (KIND == 2) fe_sqr_inl(r, r);
    if (KIND == 3) fe_sqr_f64(r, r);
    if (KIND == 4) { Fe x, y; fe_mul_inl(x, r, a); fe_sqr_inl(y, r); fe_add(r, x, y); fe_weak(r); }   // two independent ops
    if (KIND == 5) { Fe x, y; fe_mul_f64(x, r, a); fe_sqr_f64(y, r); fe_add(r, x, y); fe_weak(r); }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc ^= r.v[i] * (i + 1);
  out[t] = acc;
}

template <int KIND>
static void run(const char* name, const uint32_t* d_in, uint32_t* d_out, int sms, int iters) {
  for (int wps : {1, 2, 3, 4}) {
    int blocks = sms * wps;  // 128-thread blocks: wps blocks per SM = wps warps per scheduler
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_chain<KIND><<<blocks, 128>>>(d_in, d_out, iters / 4); CK(cudaDeviceSynchronize());
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0)); k_chain<KIND><<<blocks, 128>>>(d_in, d_out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    // clocks per op per scheduler: time * f / (iters * wps)
    double clk = best * 1e-3 * 1.965e9 / ((double)iters * wps);
    uint32_t h[4]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
    printf("{\"kind\": \"%s\", \"warps_per_scheduler\": %d, \"ms\": %.3f, \"clk_per_iter_per_scheduler\": %.1f, \"check\": \"%08x\"}\n", name, wps, best, clk, h[0] ^ h[1]);
  }
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  uint32_t *d_in, *d_out; CK(cudaMalloc(&d_in, 32 * 128 * 4)); CK(cudaMalloc(&d_out, p.multiProcessorCount * 4 * 128 * 4));
  uint32_t h[32 * 128]; srand(1); for (auto& x : h) x = (uint32_t)rand() * 2654435761u;
  CK(cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice));
  int iters = 20000;
  run<0>("mul_int", d_in, d_out, p.multiProcessorCount, iters);
  run<1>("mul_f64", d_in, d_out, p.multiProcessorCount, iters);
  run<2>("sqr_int", d_in, d_out, p.multiProcessorCount, iters);
  run<3>("sqr_f64", d_in, d_out, p.multiProcessorCount, iters);
  run<4>("mul+sqr_int", d_in, d_out, p.multiProcessorCount, iters);
  run<5>("mul+sqr_f64", d_in, d_out, p.multiProcessorCount, iters);
  return 0;
}
