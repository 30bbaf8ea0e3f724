// ed448_var.cu -- variable-base scalar multiplication [k]P (ecc/signable.rs:77 `*pub_key * h_scalar`,
// ecc/encryptable.rs:37,78): per-item table 1P..8P in local memory, signed radix-16 fixed window.
#ifndef CAPY_ED_MINBLOCKS
#define CAPY_ED_MINBLOCKS 1
#endif
#include "ed448_kernels.h"

#ifdef CAPY_VB_MINBLOCKS
#undef CAPY_ED_MINBLOCKS
#define CAPY_ED_MINBLOCKS CAPY_VB_MINBLOCKS
#endif

#ifndef CAPY_VB_BLOCK
#define CAPY_VB_BLOCK 128
#endif
// CAPY_VB_SYNC: the warps of a block pass every doubling / addition together, so they fetch the same
// instruction lines at the same time (one instruction stream per block instead of one per warp)
#if defined(CAPY_VB_SYNC) && defined(__CUDA_ARCH__)
#define CAPY_VB_BARRIER() __syncthreads()
#else
#define CAPY_VB_BARRIER() ((void)0)
#endif

namespace capy {

#ifdef CAPY_VB_HOT_INLINE
// The window loop with the field multiplications INLINED and exactly one copy of the doubling and one of the
// addition (the four doublings of a window are a rolled loop), so the loop body is ~17 multiplication bodies of
// straight-line code that the instruction prefetcher streams, while operands stay in registers instead of
// going through local memory around every out-of-line call.  Cold code (table build, validation) keeps the
// out-of-line multiplications.
__device__ __forceinline__ void vb_double_hot(PtExt& r, bool want_t) {
  Fe A, B, C, E, F, G, H, s, ZZ;
  fe_sqr_inl(A, r.X);
  fe_sqr_inl(B, r.Y);
  fe_sqr_inl(ZZ, r.Z);
  fe_add(s, r.X, r.Y);   // alpha 2
  fe_sqr_inl(E, s);      // 2 x 2
  fe_add(G, A, B);       // alpha 2
  fe_sub4(E, E, G);      // alpha 5
  fe_weak(E);            // tight
  fe_add(C, ZZ, ZZ);     // alpha 2
  fe_sub4(F, G, C);      // alpha 6
  fe_weak(F);            // tight
  fe_sub(H, A, B);       // alpha 3
  fe_mul_inl(r.X, E, F);
  fe_mul_inl(r.Y, G, H);
  fe_mul_inl(r.Z, F, G);
  if (want_t) fe_mul_inl(r.T, E, H);
}

__device__ __forceinline__ void vb_add_hot(PtExt& r, const PtCached& q) {
  Fe A, B, C, D, E0, s1, s2, E, F, G, H, AB;
  fe_mul_inl(A, r.X, q.X);
  fe_mul_inl(B, r.Y, q.Y);
  fe_mul_inl(C, r.T, q.Td);
  fe_mul_inl(D, r.Z, q.Z);
  fe_add(s1, r.X, r.Y);
  fe_add(s2, q.X, q.Y);
  fe_mul_inl(E0, s1, s2);
  fe_add(AB, A, B);     // alpha 2
  fe_sub4(E, E0, AB);   // alpha 5
  fe_weak(E);           // tight
  fe_sub(F, D, C);      // alpha 3
  fe_add(G, D, C);      // alpha 2
  fe_sub(H, B, A);      // alpha 3
  fe_mul_inl(r.X, E, F);
  fe_mul_inl(r.Y, G, H);
  fe_mul_inl(r.Z, F, G);
  fe_mul_inl(r.T, E, H);
}

__device__ __forceinline__ void pt_var_base_mul_hot(PtExt& r, const Sc& k, const PtExt& p, PtCached* tab, int8_t* dig,
                                                    bool constant_time) {
  vb_build_table(tab, p);
  sc_recode_radix16(dig, k);
  pt_identity(r);
#pragma unroll 1
  for (int i = 112; i >= 0; i--) {
    if (i != 112) {
#pragma unroll 1
      for (int d = 0; d < 4; d++) {
        CAPY_VB_BARRIER();
        vb_double_hot(r, d == 3);
      }
    }
    PtCached e;
    vb_lookup(e, tab, (int)dig[i], constant_time);
    CAPY_VB_BARRIER();
#ifdef CAPY_VB_ADD_OOL
    pt_add_cached<true>(r, r, e);  // once per window: out-of-line multiplications keep the loop body small
#else
    vb_add_hot(r, e);
#endif
  }
}
#endif

// r_i = [k_i]P_i (+ addend_i).  scalars: 56-byte big-endian, exact integers (mode4 = 0) or
// 4 * BE mod r (mode4 = 1, ECDH: ecc/encryptable.rs:36).  Off-curve P_i -> bad[i] = 1, identity out.
__global__ void __launch_bounds__(CAPY_VB_BLOCK, CAPY_ED_MINBLOCKS) var_base_kernel(const uint8_t* __restrict__ scalars_be56, int mode4,
                                                       const uint8_t* __restrict__ points_xy,
                                                       const uint32_t* __restrict__ addend /* ext SoA or null */,
                                                       uint32_t* __restrict__ proj, uint8_t* __restrict__ bad, uint64_t n,
                                                       int constant_time) {
  // no early exits: idle threads of the last block shadow the last item and off-curve inputs run the (complete)
  // arithmetic on the identity, so every thread reaches the block barriers of the window loop
  const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = gi < n;
  const uint64_t i = active ? gi : n - 1;
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
  if (!ok) pt_identity(p);
  PtCached tab[8];
  int8_t dig[113];
#ifdef CAPY_VB_HOT_INLINE
  pt_var_base_mul_hot(r, k, p, tab, dig, constant_time != 0);
#else
  pt_var_base_mul(r, k, p, tab, dig, constant_time != 0);
#endif
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
  if (!ok) pt_identity(r);  // off-curve input: identity out, no addend (the caller reports the item as bad)
  if (active) {
    if (bad) bad[i] = ok ? 0 : 1;
    store_ext(proj, n, i, r);
  }
}

int launch_var_base(capy_ctx* ctx, cudaStream_t st, const uint8_t* scalars, int mode4, const uint8_t* points,
                           const uint32_t* addend, uint32_t* proj, uint8_t* bad, uint64_t n, bool constant_time) {
  var_base_kernel<<<grid_for(n, CAPY_VB_BLOCK), CAPY_VB_BLOCK, 0, st>>>(scalars, mode4, points, addend, proj, bad, n, constant_time ? 1 : 0);
  ctx->launches++;
  CAPY_CUDA(ctx, cudaGetLastError());
  return CAPY_OK;
}


}  // namespace capy
