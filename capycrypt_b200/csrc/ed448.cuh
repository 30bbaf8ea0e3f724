// ed448.cuh -- Ed448-Goldilocks group arithmetic for sm_100a (untwisted Edwards curve
// x^2 + y^2 = 1 + d x^2 y^2, d = -39081, a = 1).
//
// Replaces `ExtendedPoint::{generator, Mul<Scalar>, Add, to_affine}` of the un-vendored crate
// tiny_ed448_goldilocks 0.1.8 (call sites ecc/keypair.rs:44, ecc/signable.rs:48-49,77,
// ecc/encryptable.rs:37-38,78).  Extended coordinates (X:Y:Z:T), complete formulas
// (add-2008-hwcd / dbl-2008-hwcd with a = 1; d is a non-square so the unified addition has no
// exceptional points), so every input -- identity, small-order points, P + P -- takes the same path.
//
// Scalar multiplication is a fixed-window ladder on signed radix-16 digits:
//   fixed base   112 windows x 8 affine entries (j+1) * 16^i * G  -> 112 mixed additions, no doubling
//   variable     per-item table 1P..8P, 113 windows x (4 doublings + 1 addition)
// Table entries are chosen by a constant-time scan (every entry is read, the wanted one is kept
// with a mask) unless the caller states that the scalar is public (verify).
#pragma once
#include "fp448.cuh"
#include "sc448.cuh"

namespace capy {

constexpr uint32_t EDW_D_ABS = 39081u;  // d = -39081

struct PtExt {  // extended projective, all coordinates tight
  Fe X, Y, Z, T;
};
struct PtAffN {  // affine table entry: x, y, td = d*x*y
  Fe x, y, td;
};
struct PtCached {  // projective table entry: X, Y, Z, Td = d*T
  Fe X, Y, Z, Td;
};

CAPY_HD void pt_identity(PtExt& p) {
  fe_zero(p.X);
  fe_one(p.Y);
  fe_one(p.Z);
  fe_zero(p.T);
}

// r = d * a = -(39081 * a)
CAPY_HD void fe_mul_d(Fe& r, const Fe& a) {
  uint64_t R[16];
#pragma unroll
  for (int i = 0; i < 16; i++) R[i] = (uint64_t)a.v[i] * EDW_D_ABS;
  Fe t;
  fe_carry_wide(t, R);
  fe_neg(r, t);
}

// shared tail of the additions: given A = X1*X2, B = Y1*Y2, C = d*T1*T2, D = Z1*Z2 (tight) and
// E0 = (X1+Y1)*(X2+Y2) (tight):  E = E0 - A - B, F = D - C, G = D + C, H = B - A  (a = 1)
template <bool WANT_T>
CAPY_HD void add_tail(PtExt& r, const Fe& A, const Fe& B, const Fe& C, const Fe& D, const Fe& E0) {
  Fe E, F, G, H, AB;
  fe_add(AB, A, B);     // alpha 2
  fe_sub4(E, E0, AB);   // alpha 5
  fe_weak(E);           // tight
  fe_sub(F, D, C);      // alpha 3
  fe_add(G, D, C);      // alpha 2
  fe_sub(H, B, A);      // alpha 3
  fe_mul(r.X, E, F);    // 1 x 3
  fe_mul(r.Y, G, H);    // 2 x 3
  fe_mul(r.Z, F, G);    // 3 x 2
  if (WANT_T) fe_mul(r.T, E, H);  // 1 x 3
}

// r = p + q, q an affine table entry (Z2 = 1): 8 M
template <bool WANT_T>
CAPY_HD void pt_madd(PtExt& r, const PtExt& p, const PtAffN& q) {
  Fe A, B, C, E0, s1, s2;
  fe_mul(A, p.X, q.x);
  fe_mul(B, p.Y, q.y);
  fe_mul(C, p.T, q.td);
  fe_add(s1, p.X, p.Y);  // alpha 2
  fe_add(s2, q.x, q.y);  // alpha 2
  fe_mul(E0, s1, s2);    // 2 x 2
  Fe D;
  fe_copy(D, p.Z);
  add_tail<WANT_T>(r, A, B, C, D, E0);
}

// r = p + q, q a projective table entry: 9 M
template <bool WANT_T>
CAPY_HD void pt_add_cached(PtExt& r, const PtExt& p, const PtCached& q) {
  Fe A, B, C, D, E0, s1, s2;
  fe_mul(A, p.X, q.X);
  fe_mul(B, p.Y, q.Y);
  fe_mul(C, p.T, q.Td);
  fe_mul(D, p.Z, q.Z);
  fe_add(s1, p.X, p.Y);
  fe_add(s2, q.X, q.Y);
  fe_mul(E0, s1, s2);
  add_tail<WANT_T>(r, A, B, C, D, E0);
}

// r = 2p: 4 S + 3 M (+1 M for T).  Does not read p.T.
template <bool WANT_T>
CAPY_HD void pt_double(PtExt& r, const PtExt& p) {
  Fe A, B, C, E, F, G, H, s, ZZ;
  fe_sqr(A, p.X);
  fe_sqr(B, p.Y);
  fe_sqr(ZZ, p.Z);
  fe_add(s, p.X, p.Y);   // alpha 2
  fe_sqr(E, s);          // 2 x 2
  fe_add(G, A, B);       // alpha 2
  fe_sub4(E, E, G);      // alpha 5
  fe_weak(E);            // tight
  fe_add(C, ZZ, ZZ);     // alpha 2
  fe_sub4(F, G, C);      // alpha 6
  fe_weak(F);            // tight
  fe_sub(H, A, B);       // alpha 3
  fe_mul(r.X, E, F);     // 1 x 1
  fe_mul(r.Y, G, H);     // 2 x 3
  fe_mul(r.Z, F, G);     // 1 x 2
  if (WANT_T) fe_mul(r.T, E, H);  // 1 x 3
}

CAPY_HD void pt_to_cached(PtCached& c, const PtExt& p) {
  fe_copy(c.X, p.X);
  fe_copy(c.Y, p.Y);
  fe_copy(c.Z, p.Z);
  fe_mul_d(c.Td, p.T);
}

// affine (x, y) -> extended; returns false when the point is not on the curve
CAPY_HD bool pt_from_affine(PtExt& p, const Fe& x, const Fe& y) {
  fe_copy(p.X, x);
  fe_copy(p.Y, y);
  fe_one(p.Z);
  fe_mul(p.T, x, y);
  // x^2 + y^2 - 1 + 39081 x^2 y^2 == 0
  Fe x2, y2, l, m, one;
  fe_sqr(x2, x);
  fe_sqr(y2, y);
  fe_mul(m, x2, y2);
  fe_mul_d(m, m);        // d x^2 y^2 (tight)
  fe_add(l, x2, y2);     // alpha 2
  fe_one(one);
  fe_add(m, m, one);     // 1 + d x^2 y^2, alpha ~1
  fe_sub(l, l, m);       // alpha 4
  return fe_is_zero(l);
}

// ---- signed radix-16 recoding -----------------------------------------------------------------
// digits d_i in [-8, 7] (i < 112) with sum d_i 16^i = k; dig[112] is the final carry (0 or 1),
// which is always 0 when k < 2^446 (reduced scalars).
CAPY_HD void sc_recode_radix16(int8_t* dig /*[113]*/, const Sc& k) {
  uint32_t carry = 0;
#pragma unroll 1
  for (int i = 0; i < 112; i++) {
    uint32_t nib = ((k.w[i >> 3] >> (4 * (i & 7))) & 15u) + carry;
    carry = nib >= 8u ? 1u : 0u;
    dig[i] = (int8_t)((int32_t)nib - (int32_t)(carry << 4));
  }
  dig[112] = (int8_t)carry;
}

// constant-time conditional negation of an affine entry: -(x, y) = (-x, y), td -> -td
CAPY_HD void ptaff_cneg(PtAffN& e, uint32_t neg_mask) {
  Fe nx, ntd;
  fe_neg(nx, e.x);
  fe_neg(ntd, e.td);
  fe_cmov(e.x, nx, neg_mask);
  fe_cmov(e.td, ntd, neg_mask);
}
CAPY_HD void ptcached_cneg(PtCached& e, uint32_t neg_mask) {
  Fe nx, ntd;
  fe_neg(nx, e.X);
  fe_neg(ntd, e.Td);
  fe_cmov(e.X, nx, neg_mask);
  fe_cmov(e.Td, ntd, neg_mask);
}

// ---- fixed-base comb ---------------------------------------------------------------------------
// table layout: entry (i, j) = (j+1) * 16^i * G as 48 u32 words (x | y | td, 16 tight limbs each)
constexpr int FB_WINDOWS = 112;
constexpr int FB_ENTRIES = 8;
constexpr int FB_ENTRY_WORDS = 48;

#if defined(__CUDA_ARCH__)
#define CAPY_LD128(p) __ldg(reinterpret_cast<const uint4*>(p))
#else
#define CAPY_LD128(p) (*reinterpret_cast<const uint4*>(p))
#endif

// e = |dgt| * 16^i * G (identity representation x = 0, y = 1, td = 0 when dgt == 0), sign applied
CAPY_HD void fb_lookup(PtAffN& e, const uint32_t* __restrict__ table, int i, int dgt, bool CONSTANT_TIME) {
  const uint32_t neg = dgt < 0 ? 0xffffffffu : 0u;
  const uint32_t mag = (uint32_t)(dgt < 0 ? -dgt : dgt);  // 0..8
  uint32_t w[FB_ENTRY_WORDS];
#pragma unroll
  for (int k = 0; k < FB_ENTRY_WORDS; k++) w[k] = 0;
  w[16] = 1;  // y = 1
  const uint32_t* row = table + (size_t)i * FB_ENTRIES * FB_ENTRY_WORDS;
  if (CONSTANT_TIME) {
#pragma unroll 1
    for (uint32_t j = 1; j <= FB_ENTRIES; j++) {
      const uint32_t m = (j == mag) ? 0xffffffffu : 0u;
      const uint32_t* ent = row + (j - 1) * FB_ENTRY_WORDS;
#pragma unroll
      for (int k = 0; k < FB_ENTRY_WORDS; k += 4) {
        const uint4 v = CAPY_LD128(ent + k);
        w[k + 0] = (v.x & m) | (w[k + 0] & ~m);
        w[k + 1] = (v.y & m) | (w[k + 1] & ~m);
        w[k + 2] = (v.z & m) | (w[k + 2] & ~m);
        w[k + 3] = (v.w & m) | (w[k + 3] & ~m);
      }
    }
  } else if (mag != 0) {
    const uint32_t* ent = row + (mag - 1) * FB_ENTRY_WORDS;
#pragma unroll
    for (int k = 0; k < FB_ENTRY_WORDS; k += 4) {
      const uint4 v = CAPY_LD128(ent + k);
      w[k + 0] = v.x;
      w[k + 1] = v.y;
      w[k + 2] = v.z;
      w[k + 3] = v.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 16; k++) {
    e.x.v[k] = w[k];
    e.y.v[k] = w[16 + k];
    e.td.v[k] = w[32 + k];
  }
  ptaff_cneg(e, neg);
}

// r = [k]G for k in [0, r) via the comb table
CAPY_HD void pt_fixed_base_mul(PtExt& r, const Sc& k, const uint32_t* __restrict__ table, bool CONSTANT_TIME) {
  pt_identity(r);
  uint32_t carry = 0;
#pragma unroll 1
  for (int i = 0; i < FB_WINDOWS; i++) {
    uint32_t nib = ((k.w[i >> 3] >> (4 * (i & 7))) & 15u) + carry;
    carry = nib >= 8u ? 1u : 0u;
    const int dgt = (int)nib - (int)(carry << 4);
    PtAffN e;
    fb_lookup(e, table, i, dgt, CONSTANT_TIME);
    pt_madd<true>(r, r, e);
  }
}

// ---- variable base ------------------------------------------------------------------------------
// tab[j] = (j+1) * P, j = 0..7, as cached projective entries
CAPY_HD void vb_build_table(PtCached* tab /*[8]*/, const PtExt& p) {
  PtExt acc;
  pt_to_cached(tab[0], p);
  pt_double<true>(acc, p);
  pt_to_cached(tab[1], acc);
#pragma unroll 1
  for (int j = 2; j < 8; j++) {
    pt_add_cached<true>(acc, acc, tab[0]);
    pt_to_cached(tab[j], acc);
  }
}

CAPY_HD void vb_lookup(PtCached& e, const PtCached* tab, int dgt, bool CONSTANT_TIME) {
  const uint32_t neg = dgt < 0 ? 0xffffffffu : 0u;
  const uint32_t mag = (uint32_t)(dgt < 0 ? -dgt : dgt);
  // identity as a cached entry: X = 0, Y = 1, Z = 1, Td = 0
  fe_zero(e.X);
  fe_one(e.Y);
  fe_one(e.Z);
  fe_zero(e.Td);
  if (CONSTANT_TIME) {
#pragma unroll 1
    for (uint32_t j = 1; j <= 8; j++) {
      const uint32_t m = (j == mag) ? 0xffffffffu : 0u;
      fe_cmov(e.X, tab[j - 1].X, m);
      fe_cmov(e.Y, tab[j - 1].Y, m);
      fe_cmov(e.Z, tab[j - 1].Z, m);
      fe_cmov(e.Td, tab[j - 1].Td, m);
    }
  } else if (mag != 0) {
    e = tab[mag - 1];
  }
  ptcached_cneg(e, neg);
}

// r = [k]P for the exact integer k < 2^448 (NOT reduced: quirk Q10, ecc/signable.rs:76-77)
CAPY_HD void pt_var_base_mul(PtExt& r, const Sc& k, const PtExt& p, PtCached* tab /*[8] scratch*/, int8_t* dig /*[113]*/,
                             bool CONSTANT_TIME) {
  vb_build_table(tab, p);
  sc_recode_radix16(dig, k);
  pt_identity(r);
#pragma unroll 1
  for (int i = 112; i >= 0; i--) {
    if (i != 112) {
      pt_double<false>(r, r);
      pt_double<false>(r, r);
      pt_double<false>(r, r);
      pt_double<true>(r, r);
    }
    PtCached e;
    vb_lookup(e, tab, (int)dig[i], CONSTANT_TIME);
    pt_add_cached<true>(r, r, e);
  }
}

}  // namespace capy

// ---- generator and comb-table construction --------------------------------------------------------
namespace capy {

// RFC 8032 Ed448 base point, 28-bit limbs
#define CAPY_GX_LIMBS                                                                                           \
  {0x70cc05eu, 0x26a82bcu, 0x0938e26u, 0x80e18b0u, 0x511433bu, 0xf72ab66u, 0x412ae1au, 0xa3d3a46u, 0xa6de324u, \
   0x0f1767eu, 0x4657047u, 0x36da9e1u, 0x5a622bfu, 0xed221d1u, 0x66bed0du, 0x4f1970cu}
#define CAPY_GY_LIMBS                                                                                           \
  {0x230fa14u, 0x08795bfu, 0x7c8ad98u, 0x132c4edu, 0x9c4fdbdu, 0x1ce67c3u, 0x73ad3ffu, 0x05a0c2du, 0x7789c1eu, \
   0xa398408u, 0xa73736cu, 0xc7624beu, 0x03756c9u, 0x2488762u, 0x16eb6bcu, 0x693f467u}

CAPY_HD void pt_generator(PtExt& g) {
  const uint32_t gx[16] = CAPY_GX_LIMBS;
  const uint32_t gy[16] = CAPY_GY_LIMBS;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    g.X.v[i] = gx[i];
    g.Y.v[i] = gy[i];
  }
  fe_one(g.Z);
  fe_mul(g.T, g.X, g.Y);
}

// writes the 8 entries of window i: (j+1) * 16^i * G, affine, canonical limbs
CAPY_HD void fb_build_window(uint32_t* row /*[8 * 48]*/, int i) {
  PtExt base;
  pt_generator(base);
#pragma unroll 1
  for (int k = 0; k < 4 * i; k++) pt_double<true>(base, base);
  PtCached cb;
  pt_to_cached(cb, base);
  PtExt acc = base;
#pragma unroll 1
  for (int j = 0; j < FB_ENTRIES; j++) {
    if (j > 0) pt_add_cached<true>(acc, acc, cb);
    Fe zi, x, y, td;
    fe_inv(zi, acc.Z);
    fe_mul(x, acc.X, zi);
    fe_mul(y, acc.Y, zi);
    fe_mul(td, x, y);
    fe_mul_d(td, td);
    fe_canon(x);
    fe_canon(y);
    fe_canon(td);
    uint32_t* e = row + j * FB_ENTRY_WORDS;
    for (int k = 0; k < 16; k++) {
      e[k] = x.v[k];
      e[16 + k] = y.v[k];
      e[32 + k] = td.v[k];
    }
  }
}

}  // namespace capy
