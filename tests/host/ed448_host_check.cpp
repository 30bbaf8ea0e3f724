// tests/host/ed448_host_check.cpp -- TEST SCAFFOLDING.  Compiles the engine's __host__ __device__
// Ed448 arithmetic (capycrypt_b200/csrc/{fp448,sc448,ed448}.cuh) for the CPU so that the exact code
// the GPU kernels run can be checked against the oracle in the GPU-less build container
// (tests/test_ed448_host.py).  It is never part of libcapycrypt_gpu.so and nothing in the
// product loads it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../capycrypt_b200/csrc/ed448.cuh"

using namespace capy;

static void fe_from_le56(Fe& f, const uint8_t* b) {
  uint32_t w[14];
  memcpy(w, b, 56);
  fe_from_words(f, w);
}
static void fe_to_le56(uint8_t* b, const Fe& f) {
  uint32_t w[14];
  fe_to_words(w, f);
  memcpy(b, w, 56);
}
static void pt_out(uint8_t* xy, const PtExt& p) {
  Fe zi, x, y;
  fe_inv(zi, p.Z);
  fe_mul(x, p.X, zi);
  fe_mul(y, p.Y, zi);
  fe_to_le56(xy, x);
  fe_to_le56(xy + 56, y);
}

static std::vector<uint32_t> g_table;
static const uint32_t* table() {
  if (g_table.empty()) {
    g_table.resize((size_t)FB_WINDOWS * FB_ENTRIES * FB_ENTRY_WORDS);
    for (int i = 0; i < FB_WINDOWS; i++) fb_build_window(g_table.data() + (size_t)i * FB_ENTRIES * FB_ENTRY_WORDS, i);
  }
  return g_table.data();
}

extern "C" {
// op: 0 mul, 1 sqr, 2 inv, 3 add-then-canon, 4 sub-then-canon, 5 mul of (a+b)*(a-b) (loose operands)
void host_fe_op(int op, const uint8_t* a56, const uint8_t* b56, uint8_t* out56) {
  Fe a, b, r;
  fe_from_le56(a, a56);
  fe_from_le56(b, b56);
  switch (op) {
    case 0: fe_mul(r, a, b); break;
    case 1: fe_sqr(r, a); break;
    case 2: fe_inv(r, a); break;
    case 3: fe_add(r, a, b); break;
    case 4: fe_sub(r, a, b); break;
    case 5: { Fe s, d; fe_add(s, a, b); fe_sub(d, a, b); fe_mul(r, s, d); break; }
    default: fe_zero(r);
  }
  fe_to_le56(out56, r);
}
// op: 0 mul_mod, 1 mul4_mod, 2 sub_mod (inputs reduced first), 3 reduce
void host_sc_op(int op, const uint8_t* a_be, const uint8_t* b_be, uint8_t* out_be) {
  Sc a, b, r;
  sc_from_be(a, a_be);
  sc_from_be(b, b_be);
  switch (op) {
    case 0: sc_mul_mod(r, a, b); break;
    case 1: sc_mul4_mod(r, a); break;
    case 2: { Sc ar, br; sc_reduce_448(ar, a); sc_reduce_448(br, b); sc_sub_mod(r, ar, br); break; }
    default: sc_reduce_448(r, a);
  }
  sc_to_be(out_be, r);
}
void host_fixed_base(const uint8_t* k_be, int constant_time, uint8_t* out_xy) {
  Sc k, kr;
  sc_from_be(k, k_be);
  sc_reduce_448(kr, k);
  PtExt r;
  pt_fixed_base_mul(r, kr, table(), constant_time != 0);
  pt_out(out_xy, r);
}
int host_var_base2(const uint8_t* k_be, const uint8_t* pt_xy, int constant_time, const uint8_t* addend_xy, uint8_t* out_xy);
int host_var_base(const uint8_t* k_be, const uint8_t* pt_xy, int constant_time, uint8_t* out_xy) {
  return host_var_base2(k_be, pt_xy, constant_time, nullptr, out_xy);
}
// [k]P (+ addend): the body of var_base_kernel with the per-item table on the heap (stride 3 between chunks, to
// exercise the interleaving arithmetic)
int host_var_base2(const uint8_t* k_be, const uint8_t* pt_xy, int constant_time, const uint8_t* addend_xy, uint8_t* out_xy) {
  Sc k;
  sc_from_be(k, k_be);
  Fe x, y;
  fe_from_le56(x, pt_xy);
  fe_from_le56(y, pt_xy + 56);
  PtExt r;
  if (!pt_from_affine(r, x, y)) return -4;
  constexpr int STRIDE = 3;
  std::vector<uint4> tab((size_t)VB_ENTRIES * VB_CHUNKS * STRIDE);
  PtExt add;
  if (addend_xy) {
    fe_from_le56(x, addend_xy);
    fe_from_le56(y, addend_xy + 56);
    if (!pt_from_affine(add, x, y)) return -4;
  }
  pt_var_base_mul<STRIDE>(r, k, tab.data() + 1, constant_time != 0, addend_xy != nullptr, [&](PtExt& a) { a = add; });
  pt_out(out_xy, r);
  return 0;
}
void host_table_entry(int i, int j, uint8_t* out_xy_td /*168: ymx | ypx | td2*/) {
  const uint32_t* e = table() + ((size_t)i * FB_ENTRIES + j) * FB_ENTRY_WORDS;
  Fe x, y, td;
  for (int k = 0; k < 16; k++) { x.v[k] = e[k]; y.v[k] = e[16 + k]; td.v[k] = e[32 + k]; }
  fe_to_le56(out_xy_td, x);
  fe_to_le56(out_xy_td + 56, y);
  fe_to_le56(out_xy_td + 112, td);
}
}

// z = k - (BE(h) * s mod r) mod r with k, s already reduced -- the body of sign_finish_kernel
extern "C" void host_sign_finish(const uint8_t* k_be, const uint8_t* s_be, const uint8_t* h56, uint8_t* z_be) {
  Sc k, s, h, hs, z;
  sc_from_be(k, k_be);
  sc_from_be(s, s_be);
  sc_from_be(h, h56);
  sc_mul_mod(hs, h, s);
  sc_sub_mod(z, k, hs);
  sc_to_be(z_be, z);
}
