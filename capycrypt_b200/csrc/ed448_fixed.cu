// ed448_fixed.cu -- fixed-base scalar multiplication [k]G: comb table of G built once per device,
// then 112 mixed additions per item (the reference has no fixed-base path: `generator() * s` runs the
// generic Mul<Scalar>, ecc/keypair.rs:44; the result is the same curve point).
#ifndef CAPY_ED_MINBLOCKS
#define CAPY_ED_MINBLOCKS 1
#endif
#include "ed448_kernels.h"

namespace capy {

struct Ed448Tables {
  uint32_t* fb = nullptr;  // [112][8][48] comb table of G
};

void ed448_tables_free(DeviceCtx& dc) {
  if (dc.ed) {
    if (dc.ed->fb) cudaFree(dc.ed->fb);
    delete dc.ed;
    dc.ed = nullptr;
  }
}

__global__ void fb_table_kernel(uint32_t* table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < FB_WINDOWS) fb_build_window(table + (size_t)i * FB_ENTRIES * FB_ENTRY_WORDS, i);
}

// r_i = [k_i]G, k_i reduced scalars in SoA words; result stored extended (X, Y, Z, T)
__global__ void __launch_bounds__(128, CAPY_ED_MINBLOCKS) fixed_base_kernel(const uint32_t* __restrict__ k_words,
                                                         const uint32_t* __restrict__ table,
                                                         uint32_t* __restrict__ proj, uint64_t n, int constant_time) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Sc k;
#pragma unroll
  for (int j = 0; j < 14; j++) k.w[j] = k_words[(uint64_t)j * n + i];
  PtExt r;
  pt_fixed_base_mul(r, k, table, constant_time != 0);
  store_ext(proj, n, i, r);
}

static int ensure_tables(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream) {
  if (dc.ed && dc.ed->fb) return CAPY_OK;
  if (!dc.ed) dc.ed = new Ed448Tables();
  const size_t bytes = (size_t)FB_WINDOWS * FB_ENTRIES * FB_ENTRY_WORDS * sizeof(uint32_t);
  CAPY_CUDA(ctx, cudaMalloc(&dc.ed->fb, bytes));
  fb_table_kernel<<<(FB_WINDOWS + 31) / 32, 32, 0, stream>>>(dc.ed->fb);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}

int launch_fixed_base(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t st, const uint32_t* k_words, uint32_t* proj,
                             uint64_t n, bool constant_time) {
  int rc = ensure_tables(ctx, dc, st);
  if (rc) return rc;
  fixed_base_kernel<<<grid_for(n, 128), 128, 0, st>>>(k_words, dc.ed->fb, proj, n, constant_time ? 1 : 0);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}


}  // namespace capy
