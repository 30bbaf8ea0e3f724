// tests/host/device_selfcheck.cu -- TEST SCAFFOLDING: runs the __host__ __device__ scalar / field helpers on
// the GPU and on the CPU with the same inputs and reports any difference (pins compiler/codegen issues).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../capycrypt_b200/csrc/ed448.cuh"
using namespace capy;

struct Case { uint8_t a[56], b[56], c[56]; };
struct Out { uint8_t mulmod[56], mul4[56], red[56], submod[56], finish[56], femul[56], fesqr[56]; uint32_t prod[28]; uint8_t red28[56]; };

__host__ __device__ void run_case(const Case& in, Out& o) {
  Sc a, b, c, r, ar, br;
  sc_from_be(a, in.a); sc_from_be(b, in.b); sc_from_be(c, in.c);
  sc_mul_mod(r, a, b); sc_to_be(o.mulmod, r);
  sc_mul4_mod(r, a); sc_to_be(o.mul4, r);
  sc_reduce_448(ar, a); sc_to_be(o.red, ar);
  sc_reduce_448(br, b);
  sc_sub_mod(r, ar, br); sc_to_be(o.submod, r);
  {  // raw schoolbook product (same loop as sc_mul_mod) and the reduction of the HOST-independent value
    uint32_t p[28];
    for (int i = 0; i < 28; i++) p[i] = 0;
#pragma unroll 1
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
    for (int i = 0; i < 28; i++) o.prod[i] = p[i];
    // reduce a fixed 28-limb pattern derived from the inputs only (no dependence on the product)
    uint32_t q[28];
    for (int i = 0; i < 14; i++) { q[i] = a.w[i]; q[14 + i] = b.w[i]; }
    Sc rr; sc_reduce<28>(rr, q); sc_to_be(o.red28, rr);
  }
  // sign_finish: z = ar - (c * br mod r)
  Sc hs, z;
  sc_mul_mod(hs, c, br);
  sc_sub_mod(z, ar, hs);
  sc_to_be(o.finish, z);
  Fe x, y, m;
  uint32_t w[14];
  memcpy(w, in.a, 56); fe_from_words(x, w);
  memcpy(w, in.b, 56); fe_from_words(y, w);
  fe_mul(m, x, y); fe_to_words(w, m); memcpy(o.femul, w, 56);
  fe_sqr(m, x); fe_to_words(w, m); memcpy(o.fesqr, w, 56);
}
__global__ void k(const Case* in, Out* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) run_case(in[i], out[i]);
}
int main() {
  const int n = 4096;
  std::vector<Case> h(n);
  srand(1);
  for (auto& c : h) for (int i = 0; i < 56; i++) { c.a[i] = rand(); c.b[i] = rand(); c.c[i] = rand(); }
  memset(h[0].a, 0xff, 56); memset(h[0].b, 0xff, 56); memset(h[0].c, 0xff, 56);
  memset(h[1].a, 0, 56); memset(h[1].b, 0, 56);
  Case* d_in; Out* d_out;
  cudaMalloc(&d_in, n * sizeof(Case)); cudaMalloc(&d_out, n * sizeof(Out));
  cudaMemcpy(d_in, h.data(), n * sizeof(Case), cudaMemcpyHostToDevice);
  k<<<(n + 63) / 64, 64>>>(d_in, d_out, n);
  std::vector<Out> g(n);
  cudaError_t e = cudaMemcpy(g.data(), d_out, n * sizeof(Out), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  int bad[9] = {0};
  for (int i = 0; i < n; i++) {
    Out c; run_case(h[i], c);
    bad[0] += memcmp(c.mulmod, g[i].mulmod, 56) != 0; bad[1] += memcmp(c.mul4, g[i].mul4, 56) != 0;
    bad[2] += memcmp(c.red, g[i].red, 56) != 0; bad[3] += memcmp(c.submod, g[i].submod, 56) != 0;
    bad[4] += memcmp(c.finish, g[i].finish, 56) != 0; bad[5] += memcmp(c.femul, g[i].femul, 56) != 0;
    bad[6] += memcmp(c.fesqr, g[i].fesqr, 56) != 0;
    bad[7] += memcmp(c.prod, g[i].prod, 112) != 0; bad[8] += memcmp(c.red28, g[i].red28, 56) != 0;
  }
  printf("{\"n\": %d, \"mismatch\": {\"mul_mod\": %d, \"mul4\": %d, \"reduce\": %d, \"sub_mod\": %d, \"sign_finish\": %d, \"fe_mul\": %d, \"fe_sqr\": %d, \"raw_product\": %d, \"reduce28\": %d}}\n",
         n, bad[0], bad[1], bad[2], bad[3], bad[4], bad[5], bad[6], bad[7], bad[8]);
  return (bad[0] | bad[1] | bad[2] | bad[3] | bad[4] | bad[5] | bad[6] | bad[7] | bad[8]) ? 1 : 0;
}
