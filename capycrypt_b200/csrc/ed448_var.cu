// ed448_var.cu -- variable-base scalar multiplication [k]P (ecc/signable.rs:77 `*pub_key * h_scalar`,
// ecc/encryptable.rs:37,78): per-item table 1P..8P in local memory, signed radix-16 fixed window.
#ifndef CAPY_ED_MINBLOCKS
#define CAPY_ED_MINBLOCKS 1
#endif
#include "ed448_kernels.h"

namespace capy {

// r_i = [k_i]P_i (+ addend_i).  scalars: 56-byte big-endian, exact integers (mode4 = 0) or
// 4 * BE mod r (mode4 = 1, ECDH: ecc/encryptable.rs:36).  Off-curve P_i -> bad[i] = 1, identity out.
__global__ void __launch_bounds__(128, CAPY_ED_MINBLOCKS) var_base_kernel(const uint8_t* __restrict__ scalars_be56, int mode4,
                                                       const uint8_t* __restrict__ points_xy,
                                                       const uint32_t* __restrict__ addend /* ext SoA or null */,
                                                       uint32_t* __restrict__ proj, uint8_t* __restrict__ bad, uint64_t n,
                                                       int constant_time) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Sc k;
  sc_from_be(k, scalars_be56 + 56 * i);
  if (mode4) {
    Sc t = k;
    sc_mul4_mod(k, t);
  }
  uint32_t w[14];
  Fe x, y;
  const uint8_t* pb = points_xy + 112 * i;
#pragma unroll
  for (int j = 0; j < 14; j++) w[j] = (uint32_t)pb[4 * j] | ((uint32_t)pb[4 * j + 1] << 8) | ((uint32_t)pb[4 * j + 2] << 16) | ((uint32_t)pb[4 * j + 3] << 24);
  fe_from_words(x, w);
#pragma unroll
  for (int j = 0; j < 14; j++) w[j] = (uint32_t)pb[56 + 4 * j] | ((uint32_t)pb[57 + 4 * j] << 8) | ((uint32_t)pb[58 + 4 * j] << 16) | ((uint32_t)pb[59 + 4 * j] << 24);
  fe_from_words(y, w);
  PtExt p, r;
  const bool ok = pt_from_affine(p, x, y);
  if (bad) bad[i] = ok ? 0 : 1;
  if (!ok) {
    pt_identity(r);
    store_ext(proj, n, i, r);
    return;
  }
  PtCached tab[8];
  int8_t dig[113];
  pt_var_base_mul(r, k, p, tab, dig, constant_time != 0);
  if (addend) {
    PtExt a;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      a.X.v[j] = addend[(uint64_t)(j)*n + i];
      a.Y.v[j] = addend[(uint64_t)(16 + j) * n + i];
      a.Z.v[j] = addend[(uint64_t)(32 + j) * n + i];
      a.T.v[j] = addend[(uint64_t)(48 + j) * n + i];
    }
    PtCached c;
    pt_to_cached(c, a);
    pt_add_cached<false>(r, r, c);
  }
  store_ext(proj, n, i, r);
}

int launch_var_base(capy_ctx* ctx, cudaStream_t st, const uint8_t* scalars, int mode4, const uint8_t* points,
                           const uint32_t* addend, uint32_t* proj, uint8_t* bad, uint64_t n, bool constant_time) {
  var_base_kernel<<<grid_for(n, 128), 128, 0, st>>>(scalars, mode4, points, addend, proj, bad, n, constant_time ? 1 : 0);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}


}  // namespace capy
