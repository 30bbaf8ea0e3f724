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

// r_i = [k_i]G, k_i reduced scalars in SoA words; result stored extended (X, Y, Z, T).
// All threads of a block walk the windows in lock step, so the 3 KB table row of window i+1 is copied to
// shared memory with cp.async while window i is being added (double buffer, one barrier per window); the
// constant-time scan then reads broadcast shared-memory words instead of waiting on L1/L2.
__global__ void __launch_bounds__(128, CAPY_ED_MINBLOCKS) fixed_base_kernel(const uint32_t* __restrict__ k_words,
                                                                            const uint32_t* __restrict__ table,
                                                                            uint32_t* __restrict__ proj, uint64_t n,
                                                                            int constant_time) {
  constexpr int ROW_WORDS = FB_ENTRIES * FB_ENTRY_WORDS;  // 768 words = 192 x 16 B
  __shared__ __align__(16) uint32_t rows[2][ROW_WORDS];
  const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = gi < n;
  const uint64_t i = active ? gi : n - 1;  // idle threads of the last block shadow the last item (barriers)
  auto stage = [&](int buf, int win) {
    const uint32_t* src = table + (size_t)win * ROW_WORDS;
    for (int c = threadIdx.x; c < ROW_WORDS / 4; c += blockDim.x) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&rows[buf][4 * c]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + 4 * c));
    }
    asm volatile("cp.async.commit_group;");
  };
  stage(0, 0);
  Sc k, kq;
#pragma unroll
  for (int j = 0; j < 14; j++) k.w[j] = k_words[(uint64_t)j * n + i];
  const Sc inv4 = {CAPY_INV4_LIMBS};
  sc_mul_mod(kq, k, inv4);  // k / 4 mod r: the comb runs on the 4-isogenous curve
  PtExt acc;
  pt_identity(acc);
  uint32_t carry = 0;
#pragma unroll 1
  for (int w = 0; w < FB_WINDOWS; w++) {
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();  // row w is visible to everyone; everyone has finished reading row w - 1
    if (w + 1 < FB_WINDOWS) stage((w + 1) & 1, w + 1);
    const int dgt = fb_digit(kq, w, carry);
    PtNiels e;
    fb_lookup_row<true>(e, rows[w & 1], dgt, constant_time != 0);
    pt_madd_tw<true>(acc, acc, e);
  }
  PtExt r;
  pt_dual_isogeny(r, acc);
  if (active) store_ext(proj, n, i, r);
}

static int ensure_tables(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream) {
  if (dc.ed && dc.ed->fb) return CAPY_OK;
  if (!dc.ed) dc.ed = new Ed448Tables();
  const size_t bytes = (size_t)FB_WINDOWS * FB_ENTRIES * FB_ENTRY_WORDS * sizeof(uint32_t);
  uint32_t* fb = nullptr;
  CAPY_CUDA(ctx, cudaMalloc(&fb, bytes));
  dc.ed->fb = fb;
  fb_table_kernel<<<(FB_WINDOWS + 31) / 32, 32, 0, stream>>>(dc.ed->fb);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  // once per device: later calls may launch on other (non-blocking) streams and must not see a half-built table
  CAPY_CUDA(ctx, cudaStreamSynchronize(stream));
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
