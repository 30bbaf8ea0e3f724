/*
 * oracle/ref_cpu.c -- CPU ORACLE AND CPU BASELINE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the library built from this file.  capycrypt_b200/ never links or loads it.
 *
 * Structure-faithful C restatement of capyCRYPT 0.7.5's two hot paths (the reference is
 * Rust and cannot be compiled in this image: no rustc/cargo).  Paths cited below are
 * relative to /root/reference.
 *
 *  SHA3 side  -- src/sha3/keccakf.rs:8-423, src/sha3/sponge.rs:10-95,
 *                src/sha3/shake_functions.rs:24-89, src/sha3/aux_functions.rs:11-68,
 *                src/sha3/constants.rs:3,38-45,62-66, src/lib.rs:137-144.
 *                Same work per call as the reference: whole-message copy into a growable
 *                buffer, suffix + pad in place, per-block temp state XORed over all 25
 *                lanes, squeeze that permutes after every block (one wasted permutation).
 *                PINNED by the reference's own KATs (tests/golden/sha3_kat.json) and by
 *                equality with oracle/ref_sha3.py on seeded length sweeps.
 *  Ed448 side -- protocol of src/ecc/keypair.rs:41-51, src/ecc/signable.rs:40-86,
 *                src/ecc/encryptable.rs:36-38,76-78.  The arithmetic is in the un-vendored
 *                crate tiny_ed448_goldilocks 0.1.8 (Cargo.lock:857-868); restated from the
 *                published maths in the crate's representation class: GF(2^448-2^224-1) as
 *                8 x 56-bit limbs with 128-bit products (fiat-crypto p448_solinas_64
 *                layout), extended Edwards points, signed radix-16 fixed-window scalar
 *                multiplication with an 8-entry table, generator * s through the SAME
 *                generic routine (the reference has no fixed-base fast path,
 *                src/ecc/keypair.rs:44).  PARITY UNPINNED against the crate (no Ed448 KAT
 *                exists in the reference); pinned against oracle/ref_ed448.py, which is
 *                itself pinned against OpenSSL Ed448/X448.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ====================================================================================
 * Keccak-f[1600]  (src/sha3/keccakf.rs:8-423)
 * ==================================================================================== */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

#define ROL(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

/* one round from state A (25 named lanes in array in[]) to out[]; lane index x + 5y */
static inline void keccak_round(const uint64_t *in, uint64_t *out, uint64_t rc) {
  uint64_t c0 = in[0] ^ in[5] ^ in[10] ^ in[15] ^ in[20];
  uint64_t c1 = in[1] ^ in[6] ^ in[11] ^ in[16] ^ in[21];
  uint64_t c2 = in[2] ^ in[7] ^ in[12] ^ in[17] ^ in[22];
  uint64_t c3 = in[3] ^ in[8] ^ in[13] ^ in[18] ^ in[23];
  uint64_t c4 = in[4] ^ in[9] ^ in[14] ^ in[19] ^ in[24];
  uint64_t d0 = c4 ^ ROL(c1, 1), d1 = c0 ^ ROL(c2, 1), d2 = c1 ^ ROL(c3, 1);
  uint64_t d3 = c2 ^ ROL(c4, 1), d4 = c3 ^ ROL(c0, 1);
  uint64_t b0, b1, b2, b3, b4;
#define CHI(o)                     \
  out[o + 0] = b0 ^ (~b1 & b2);    \
  out[o + 1] = b1 ^ (~b2 & b3);    \
  out[o + 2] = b2 ^ (~b3 & b4);    \
  out[o + 3] = b3 ^ (~b4 & b0);    \
  out[o + 4] = b4 ^ (~b0 & b1);
  b0 = in[0] ^ d0;            b1 = ROL(in[6] ^ d1, 44);  b2 = ROL(in[12] ^ d2, 43);
  b3 = ROL(in[18] ^ d3, 21);  b4 = ROL(in[24] ^ d4, 14);
  CHI(0) out[0] ^= rc;
  b0 = ROL(in[3] ^ d3, 28);   b1 = ROL(in[9] ^ d4, 20);  b2 = ROL(in[10] ^ d0, 3);
  b3 = ROL(in[16] ^ d1, 45);  b4 = ROL(in[22] ^ d2, 61);
  CHI(5)
  b0 = ROL(in[1] ^ d1, 1);    b1 = ROL(in[7] ^ d2, 6);   b2 = ROL(in[13] ^ d3, 25);
  b3 = ROL(in[19] ^ d4, 8);   b4 = ROL(in[20] ^ d0, 18);
  CHI(10)
  b0 = ROL(in[4] ^ d4, 27);   b1 = ROL(in[5] ^ d0, 36);  b2 = ROL(in[11] ^ d1, 10);
  b3 = ROL(in[17] ^ d2, 15);  b4 = ROL(in[23] ^ d3, 56);
  CHI(15)
  b0 = ROL(in[2] ^ d2, 62);   b1 = ROL(in[8] ^ d3, 55);  b2 = ROL(in[14] ^ d4, 39);
  b3 = ROL(in[15] ^ d0, 41);  b4 = ROL(in[21] ^ d1, 2);
  CHI(20)
#undef CHI
}

void capy_ref_keccakf(uint64_t a[25]) {
  uint64_t e[25];
  for (int r = 0; r < 24; r += 2) {
    keccak_round(a, e, KRC[r]);
    keccak_round(e, a, KRC[r + 1]);
  }
}

/* ====================================================================================
 * growable byte vector -- stands in for Vec<u8> so the copy / realloc cost is the same
 * ==================================================================================== */
typedef struct { uint8_t *p; size_t len, cap; } bvec;
static void bv_reserve(bvec *v, size_t extra) {
  if (v->len + extra <= v->cap) return;
  size_t nc = v->cap ? v->cap * 2 : 64;
  while (nc < v->len + extra) nc *= 2;
  v->p = (uint8_t *)realloc(v->p, nc);
  v->cap = nc;
}
static void bv_push(bvec *v, const void *src, size_t n) {
  bv_reserve(v, n);
  if (n) memcpy(v->p + v->len, src, n);
  v->len += n;
}
static void bv_zeros(bvec *v, size_t n) {
  bv_reserve(v, n);
  memset(v->p + v->len, 0, n);
  v->len += n;
}
static void bv_byte(bvec *v, uint8_t b) { bv_push(v, &b, 1); }
static void bv_free(bvec *v) { free(v->p); v->p = NULL; v->len = v->cap = 0; }

/* ====================================================================================
 * sponge (src/sha3/sponge.rs)
 * ==================================================================================== */
static void pad_ten_one(bvec *m, size_t r) { /* sponge.rs:89-95 */
  size_t q = r - m->len % r;
  bv_zeros(m, q);
  m->p[m->len - 1] = 0x80;
}

static void bytes_to_state(const uint8_t *in, size_t len, size_t r, uint64_t s[25]) { /* :47-60 */
  size_t off = 0, lanes = (r * 8) / 64;
  memset(s, 0, 200);
  for (size_t b = 0; b < len / r; b++) {
    uint64_t st[25] = {0};
    for (size_t i = 0; i < lanes; i++) { /* bytes_to_word :63-69 */
      uint64_t lane = 0;
      for (int k = 0; k < 8; k++) lane += (uint64_t)in[off + k] << (8 * k);
      st[i] = lane;
      off += 8;
    }
    for (int i = 0; i < 25; i++) s[i] ^= st[i]; /* xor_states :81-85 */
    capy_ref_keccakf(s);
  }
}

static void sponge_absorb(bvec *m, size_t capacity_bits, uint64_t s[25]) { /* :10-17 */
  size_t r = (1600 - capacity_bits) / 8;
  if (m->len % r != 0) pad_ten_one(m, r);
  bytes_to_state(m->p, m->len, r, s);
}

/* sponge.rs:25-34; lean=1 skips the permutation whose output would be dropped */
static void sponge_squeeze(uint64_t s[25], size_t bit_length, size_t rate_bits, uint8_t *out, int lean) {
  size_t block = rate_bits / 64, produced = 0, want = bit_length / 8;
  while (produced * 8 < bit_length) {
    for (size_t i = 0; i < block; i++)
      for (int k = 0; k < 8; k++) {
        if (produced < want) out[produced] = (uint8_t)(s[i] >> (8 * k));
        produced++;
      }
    if (!lean || produced * 8 < bit_length) capy_ref_keccakf(s);
  }
}

/* ====================================================================================
 * SP 800-185 encoders (src/sha3/aux_functions.rs:11-68)
 * ==================================================================================== */
static size_t left_encode(uint64_t v, uint8_t out[9]) { /* :34-49 */
  if (v == 0) { out[0] = 1; out[1] = 0; return 2; }
  uint8_t be[8]; int n = 0;
  for (int i = 0; i < 8; i++) be[i] = (uint8_t)(v >> (56 - 8 * i));
  int lead = 0; while (lead < 8 && be[lead] == 0) lead++;
  n = 8 - lead;
  out[0] = (uint8_t)n;
  memcpy(out + 1, be + lead, n);
  return (size_t)n + 1;
}
static void encode_string(bvec *v, const uint8_t *s, size_t n) { /* :24-28 */
  uint8_t le[9]; size_t k = left_encode((uint64_t)n * 8, le);
  bv_push(v, le, k); bv_push(v, s, n);
}
/* byte_pad :11-18 -- `content` is appended after left_encode(w); always pads w - len%w */
static void byte_pad(bvec *out, const bvec *content, size_t w) {
  uint8_t le[9]; size_t k = left_encode(w, le);
  bv_push(out, le, k);
  bv_push(out, content->p, content->len);
  bv_zeros(out, w - out->len % w);
}
static size_t bytepad_value(int d) { return d == 224 ? 172 : d == 256 ? 168 : d == 384 ? 152 : 136; } /* lib.rs:137-144 */
static size_t capacity_from_bit_length(size_t d) { /* constants.rs:38-45 */
  size_t x = d * 2; return x <= 448 ? 448 : x <= 512 ? 512 : x <= 768 ? 768 : 1024;
}
static int valid_d(int d) { return d == 224 || d == 256 || d == 384 || d == 512; }

/* ====================================================================================
 * shake / cshake / kmac_xof (src/sha3/shake_functions.rs:24-89)
 * ==================================================================================== */
static void shake_inplace(bvec *n, int d, uint8_t *out, int lean) { /* :24-32 */
  size_t to_pad = 136 - n->len % 136; /* RATE_IN_BYTES hard-coded, quirk Q2 */
  bv_byte(n, to_pad == 1 ? 0x86 : 0x06);
  uint64_t s[25];
  sponge_absorb(n, capacity_from_bit_length((size_t)d), s);
  sponge_squeeze(s, (size_t)d, 1600 - (size_t)d, out, lean);
}

int capy_ref_sha3(const uint8_t *msg, size_t len, int d, uint8_t *out, int lean) {
  if (!valid_d(d)) return -1;
  bvec m = {0};
  bv_push(&m, msg, len); /* Message owns its Vec; hashing appends to it (quirk Q5) */
  shake_inplace(&m, d, out, lean);
  bv_free(&m);
  return 0;
}

int capy_ref_cshake(const uint8_t *x, size_t xlen, size_t l_bits, const uint8_t *n, size_t nlen,
                    const uint8_t *s, size_t slen, int d, uint8_t *out, int lean) { /* :49-64 */
  if (!valid_d(d)) return -1;
  bvec enc = {0}, buf = {0};
  encode_string(&enc, n, nlen);
  encode_string(&enc, s, slen);
  byte_pad(&buf, &enc, bytepad_value(d));
  bv_push(&buf, x, xlen);
  bv_byte(&buf, 0x04);
  if (nlen == 0 && slen == 0) { /* quirk Q4: result discarded, buffer stays mutated */
    uint8_t scratch[64];
    shake_inplace(&buf, d, scratch, 0);
  }
  uint64_t st[25];
  sponge_absorb(&buf, (size_t)d, st); /* capacity = d bits, quirk Q7 */
  sponge_squeeze(st, l_bits, 1600 - (size_t)d, out, lean);
  bv_free(&enc); bv_free(&buf);
  return 0;
}

int capy_ref_kmac_xof(const uint8_t *k, size_t klen, const uint8_t *x, size_t xlen, size_t l_bits,
                      const uint8_t *s, size_t slen, int d, uint8_t *out, int lean) { /* :79-89 */
  if (!valid_d(d)) return -1;
  bvec enc = {0}, bp = {0};
  encode_string(&enc, k, klen);
  byte_pad(&bp, &enc, bytepad_value(d));
  bv_push(&bp, x, xlen);
  static const uint8_t re0[2] = {0, 1}; /* right_encode(0), aux_functions.rs:56-58 */
  bv_push(&bp, re0, 2);
  int rc = capy_ref_cshake(bp.p, bp.len, l_bits, (const uint8_t *)"KMAC", 4, s, slen, d, out, lean);
  bv_free(&enc); bv_free(&bp);
  return rc;
}

/* ---- batch loops: threads<=1 = what the reference does (serial); >1 = rayon-style ---- */
static int clamp_threads(int t) {
#ifdef _OPENMP
  if (t <= 0) t = omp_get_max_threads();
  return t;
#else
  (void)t; return 1;
#endif
}
int capy_ref_max_threads(void) { return clamp_threads(0); }

int capy_ref_sha3_batch(const uint8_t *data, const uint64_t *off, uint64_t n, int d, uint8_t *out,
                        int threads, int lean) {
  if (!valid_d(d)) return -1;
  int t = clamp_threads(threads);
  size_t ob = (size_t)d / 8;
#pragma omp parallel for num_threads(t) schedule(dynamic, 64) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++)
    capy_ref_sha3(data + off[i], (size_t)(off[i + 1] - off[i]), d, out + (size_t)i * ob, lean);
  return 0;
}

int capy_ref_cshake_batch(const uint8_t *data, const uint64_t *off, uint64_t n, size_t l_bits,
                          const uint8_t *nn, size_t nlen, const uint8_t *s, size_t slen, int d,
                          uint8_t *out, int threads, int lean) {
  if (!valid_d(d)) return -1;
  int t = clamp_threads(threads);
  size_t ob = l_bits / 8;
#pragma omp parallel for num_threads(t) schedule(dynamic, 16) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++)
    capy_ref_cshake(data + off[i], (size_t)(off[i + 1] - off[i]), l_bits, nn, nlen, s, slen, d,
                    out + (size_t)i * ob, lean);
  return 0;
}

int capy_ref_kmac_xof_batch(const uint8_t *keys, const uint64_t *koff, const uint8_t *data,
                            const uint64_t *off, uint64_t n, size_t l_bits, const uint8_t *s,
                            size_t slen, int d, uint8_t *out, int threads, int lean) {
  if (!valid_d(d)) return -1;
  int t = clamp_threads(threads);
  size_t ob = l_bits / 8;
#pragma omp parallel for num_threads(t) schedule(dynamic, 16) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++)
    capy_ref_kmac_xof(keys + koff[i], (size_t)(koff[i + 1] - koff[i]), data + off[i],
                      (size_t)(off[i + 1] - off[i]), l_bits, s, slen, d, out + (size_t)i * ob, lean);
  return 0;
}

/* ====================================================================================
 * GF(p), p = 2^448 - 2^224 - 1, 8 x 56-bit limbs (fiat-crypto p448_solinas_64 layout)
 * ==================================================================================== */
#define M56 ((uint64_t)0x00FFFFFFFFFFFFFFULL)
typedef struct { uint64_t l[8]; } fe;

static void fe_weak(fe *a) {
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { a->l[i] += c; c = a->l[i] >> 56; a->l[i] &= M56; }
  a->l[0] += c; a->l[4] += c; /* 2^448 = 2^224 + 1 */
}
static void fe_add(fe *o, const fe *a, const fe *b) {
  for (int i = 0; i < 8; i++) o->l[i] = a->l[i] + b->l[i];
  fe_weak(o);
}
static void fe_sub(fe *o, const fe *a, const fe *b) {
  /* add 2p limb-wise so no limb goes negative */
  for (int i = 0; i < 8; i++) o->l[i] = a->l[i] + 2 * (i == 4 ? M56 - 1 : M56) - b->l[i];
  fe_weak(o);
}
static void fe_carry_wide(fe *o, u128 c[15]) {
  for (int k = 14; k >= 8; k--) { c[k - 8] += c[k]; c[k - 4] += c[k]; }
  u128 cy = 0;
  for (int i = 0; i < 8; i++) { c[i] += cy; o->l[i] = (uint64_t)c[i] & M56; cy = c[i] >> 56; }
  /* cy * 2^448 = cy * (2^224 + 1) */
  u128 t0 = (u128)o->l[0] + cy;
  o->l[0] = (uint64_t)t0 & M56;
  u128 t1 = (u128)o->l[1] + (t0 >> 56);
  o->l[1] = (uint64_t)t1 & M56;
  o->l[2] += (uint64_t)(t1 >> 56);
  u128 t4 = (u128)o->l[4] + cy;
  o->l[4] = (uint64_t)t4 & M56;
  u128 t5 = (u128)o->l[5] + (t4 >> 56);
  o->l[5] = (uint64_t)t5 & M56;
  o->l[6] += (uint64_t)(t5 >> 56);
}
static void fe_mul(fe *o, const fe *a, const fe *b) {
  u128 c[15];
  memset(c, 0, sizeof c);
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) c[i + j] += (u128)a->l[i] * b->l[j];
  fe_carry_wide(o, c);
}
static void fe_sqr(fe *o, const fe *a) {
  u128 c[15];
  memset(c, 0, sizeof c);
  for (int i = 0; i < 8; i++) {
    c[2 * i] += (u128)a->l[i] * a->l[i];
    for (int j = i + 1; j < 8; j++) c[i + j] += (u128)(2 * a->l[i]) * a->l[j];
  }
  fe_carry_wide(o, c);
}
static void fe_mul_small(fe *o, const fe *a, uint64_t k) {
  u128 cy = 0;
  for (int i = 0; i < 8; i++) { u128 t = (u128)a->l[i] * k + cy; o->l[i] = (uint64_t)t & M56; cy = t >> 56; }
  o->l[0] += (uint64_t)cy; o->l[4] += (uint64_t)cy;
  fe_weak(o);
}
static void fe_sqrn(fe *o, const fe *a, int n) { fe_sqr(o, a); for (int i = 1; i < n; i++) fe_sqr(o, o); }
static void fe_inv(fe *o, const fe *x) { /* x^(p-2), p-2 = [223 ones][0][222 ones][0][1] */
  fe t2, t3, t6, t12, t24, t27, t54, t108, t111, t222, t223, a;
  fe_sqr(&a, x);            fe_mul(&t2, &a, x);
  fe_sqr(&a, &t2);          fe_mul(&t3, &a, x);
  fe_sqrn(&a, &t3, 3);      fe_mul(&t6, &a, &t3);
  fe_sqrn(&a, &t6, 6);      fe_mul(&t12, &a, &t6);
  fe_sqrn(&a, &t12, 12);    fe_mul(&t24, &a, &t12);
  fe_sqrn(&a, &t24, 3);     fe_mul(&t27, &a, &t3);
  fe_sqrn(&a, &t27, 27);    fe_mul(&t54, &a, &t27);
  fe_sqrn(&a, &t54, 54);    fe_mul(&t108, &a, &t54);
  fe_sqrn(&a, &t108, 3);    fe_mul(&t111, &a, &t3);
  fe_sqrn(&a, &t111, 111);  fe_mul(&t222, &a, &t111);
  fe_sqr(&a, &t222);        fe_mul(&t223, &a, x);
  fe_sqrn(&a, &t223, 223);  fe_mul(&a, &a, &t222);
  fe_sqrn(&a, &a, 2);       fe_mul(o, &a, x);
}
static void fe_canon(fe *a) {
  int over;
  do { /* full carry propagation: terminates in <= 3 passes for any weakly reduced input */
    fe_weak(a);
    over = 0;
    for (int i = 0; i < 8; i++) over |= (a->l[i] > M56);
  } while (over);
  /* now every limb < 2^56, so a < 2^448 < 2p.  a >= p iff a + 2^224 + 1 carries out of 2^448 */
  uint64_t t[8], c = 1;
  for (int i = 0; i < 8; i++) { uint64_t v = a->l[i] + c + (i == 4 ? 1 : 0); t[i] = v & M56; c = v >> 56; }
  if (c) for (int i = 0; i < 8; i++) a->l[i] = t[i];
}
static void fe_to_bytes(uint8_t out[56], const fe *a) {
  fe t = *a; fe_canon(&t);
  for (int i = 0; i < 8; i++) for (int k = 0; k < 7; k++) out[7 * i + k] = (uint8_t)(t.l[i] >> (8 * k));
}
static void fe_from_bytes(fe *o, const uint8_t in[56]) {
  for (int i = 0; i < 8; i++) { uint64_t v = 0; for (int k = 0; k < 7; k++) v |= (uint64_t)in[7 * i + k] << (8 * k); o->l[i] = v; }
}
static int fe_is_zero(const fe *a) { uint8_t b[56]; fe_to_bytes(b, a); uint8_t r = 0; for (int i = 0; i < 56; i++) r |= b[i]; return r == 0; }

/* ====================================================================================
 * extended Edwards points, a = 1, d = -39081
 * ==================================================================================== */
typedef struct { fe X, Y, Z, T; } pt;
static const fe FE_ONE = {{1, 0, 0, 0, 0, 0, 0, 0}};
static const fe FE_ZERO = {{0, 0, 0, 0, 0, 0, 0, 0}};
#define EDW_D 39081u

static void pt_identity(pt *p) { p->X = FE_ZERO; p->Y = FE_ONE; p->Z = FE_ONE; p->T = FE_ZERO; }
static void pt_add(pt *o, const pt *p, const pt *q) {
  fe A, B, C, Dd, E, F, G, H, t0, t1;
  fe_mul(&A, &p->X, &q->X);
  fe_mul(&B, &p->Y, &q->Y);
  fe_mul(&t0, &p->T, &q->T);
  fe_mul_small(&t1, &t0, EDW_D);       /* 39081 * T1T2 ; C = -that */
  fe_sub(&C, &FE_ZERO, &t1);
  fe_mul(&Dd, &p->Z, &q->Z);
  fe_add(&t0, &p->X, &p->Y);
  fe_add(&t1, &q->X, &q->Y);
  fe_mul(&E, &t0, &t1);
  fe_sub(&E, &E, &A);
  fe_sub(&E, &E, &B);
  fe_sub(&F, &Dd, &C);
  fe_add(&G, &Dd, &C);
  fe_sub(&H, &B, &A);
  fe_mul(&o->X, &E, &F);
  fe_mul(&o->Y, &G, &H);
  fe_mul(&o->T, &E, &H);
  fe_mul(&o->Z, &F, &G);
}
static void pt_double(pt *o, const pt *p) {
  fe A, B, C, E, F, G, H, t0;
  fe_sqr(&A, &p->X);
  fe_sqr(&B, &p->Y);
  fe_sqr(&t0, &p->Z);
  fe_add(&C, &t0, &t0);
  fe_add(&t0, &p->X, &p->Y);
  fe_sqr(&E, &t0);
  fe_sub(&E, &E, &A);
  fe_sub(&E, &E, &B);
  fe_add(&G, &A, &B);
  fe_sub(&F, &G, &C);
  fe_sub(&H, &A, &B);
  fe_mul(&o->X, &E, &F);
  fe_mul(&o->Y, &G, &H);
  fe_mul(&o->T, &E, &H);
  fe_mul(&o->Z, &F, &G);
}
static void pt_cneg(pt *p, uint64_t neg) { /* constant-time conditional negate */
  fe nx, nt;
  fe_sub(&nx, &FE_ZERO, &p->X);
  fe_sub(&nt, &FE_ZERO, &p->T);
  uint64_t m = (uint64_t)0 - neg;
  for (int i = 0; i < 8; i++) {
    p->X.l[i] ^= m & (p->X.l[i] ^ nx.l[i]);
    p->T.l[i] ^= m & (p->T.l[i] ^ nt.l[i]);
  }
}
static void pt_to_affine_bytes(uint8_t out[112], const pt *p) {
  fe zi, x, y;
  fe_inv(&zi, &p->Z);
  fe_mul(&x, &p->X, &zi);
  fe_mul(&y, &p->Y, &zi);
  fe_to_bytes(out, &x);
  fe_to_bytes(out + 56, &y);
}
static int pt_from_affine_bytes(pt *p, const uint8_t in[112]) {
  fe_from_bytes(&p->X, in);
  fe_from_bytes(&p->Y, in + 56);
  p->Z = FE_ONE;
  fe_mul(&p->T, &p->X, &p->Y);
  /* on-curve: x^2 + y^2 - 1 + 39081 x^2 y^2 == 0 */
  fe x2, y2, l, r;
  fe_sqr(&x2, &p->X); fe_sqr(&y2, &p->Y);
  fe_add(&l, &x2, &y2);
  fe_sub(&l, &l, &FE_ONE);
  fe_mul(&r, &x2, &y2);
  fe_mul_small(&r, &r, EDW_D);
  fe_add(&l, &l, &r);
  return fe_is_zero(&l) ? 0 : -4;
}

/* RFC 8032 base point, 56-byte little-endian x then y */
static const uint8_t GEN_XY[112] = {
    0x5e, 0xc0, 0x0c, 0xc7, 0x2b, 0xa8, 0x26, 0x26, 0x8e, 0x93, 0x00, 0x8b, 0xe1, 0x80, 0x3b, 0x43,
    0x11, 0x65, 0xb6, 0x2a, 0xf7, 0x1a, 0xae, 0x12, 0x64, 0xa4, 0xd3, 0xa3, 0x24, 0xe3, 0x6d, 0xea,
    0x67, 0x17, 0x0f, 0x47, 0x70, 0x65, 0x14, 0x9e, 0xda, 0x36, 0xbf, 0x22, 0xa6, 0x15, 0x1d, 0x22,
    0xed, 0x0d, 0xed, 0x6b, 0xc6, 0x70, 0x19, 0x4f,
    0x14, 0xfa, 0x30, 0xf2, 0x5b, 0x79, 0x08, 0x98, 0xad, 0xc8, 0xd7, 0x4e, 0x2c, 0x13, 0xbd, 0xfd,
    0xc4, 0x39, 0x7c, 0xe6, 0x1c, 0xff, 0xd3, 0x3a, 0xd7, 0xc2, 0xa0, 0x05, 0x1e, 0x9c, 0x78, 0x87,
    0x40, 0x98, 0xa3, 0x6c, 0x73, 0x73, 0xea, 0x4b, 0x62, 0xc7, 0xc9, 0x56, 0x37, 0x20, 0x76, 0x88,
    0x24, 0xbc, 0xb6, 0x6e, 0x71, 0x46, 0x3f, 0x69};

/* `ExtendedPoint * Scalar`: signed radix-16 fixed window, 8-entry table, scalar given as the
   exact 448-bit integer (56 bytes big-endian, NOT reduced -- quirk Q10) */
static void pt_scalar_mult(pt *o, const pt *base, const uint8_t k_be[56]) {
  pt tab[8];
  tab[0] = *base;
  pt_double(&tab[1], base);
  for (int j = 2; j < 8; j++) pt_add(&tab[j], &tab[j - 1], base);
  int8_t dig[113];
  int carry = 0;
  for (int i = 0; i < 112; i++) {
    uint8_t byte = k_be[55 - i / 2];
    int nib = ((i & 1) ? (byte >> 4) : (byte & 15)) + carry;
    carry = nib > 8;
    dig[i] = (int8_t)(nib - 16 * carry);
  }
  dig[112] = (int8_t)carry;
  pt acc; pt_identity(&acc);
  for (int i = 112; i >= 0; i--) {
    if (i != 112) for (int k = 0; k < 4; k++) pt_double(&acc, &acc);
    int dgt = dig[i];
    uint64_t neg = (uint64_t)(dgt < 0);
    unsigned mag = (unsigned)(dgt < 0 ? -dgt : dgt);
    /* constant-time table scan; magnitude 0 selects the identity */
    pt sel; pt_identity(&sel);
    for (unsigned j = 1; j <= 8; j++) {
      uint64_t m = (uint64_t)0 - (uint64_t)(j == mag);
      const uint64_t *src = (const uint64_t *)&tab[j - 1];
      uint64_t *dst = (uint64_t *)&sel;
      for (int w = 0; w < 32; w++) dst[w] ^= m & (dst[w] ^ src[w]);
    }
    pt_cneg(&sel, neg);
    pt_add(&acc, &acc, &sel);
  }
  *o = acc;
}

/* ====================================================================================
 * scalars mod r  (U448 big-endian I/O, src/sha3/aux_functions.rs:102-110)
 * little-endian 32-bit limbs internally
 * ==================================================================================== */
#define SL 14
static const uint32_t R_LIMBS[SL] = {0xab5844f3u, 0x2378c292u, 0x8dc58f55u, 0x216cc272u, 0xaed63690u, 0xc44edb49u, 0x7cca23e9u,
                                     0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
/* c = 2^446 - r (224 bits) */
static const uint32_t C_LIMBS[7] = {0x54a7bb0du, 0xdc873d6du, 0x723a70aau, 0xde933d8du, 0x5129c96fu, 0x3bb124b6u, 0x8335dc16u};

static void sc_from_be(uint32_t o[SL], const uint8_t b[56]) {
  for (int i = 0; i < SL; i++) {
    const uint8_t *p = b + 52 - 4 * i;
    o[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
}
static void sc_to_be(uint8_t b[56], const uint32_t a[SL]) {
  for (int i = 0; i < SL; i++) {
    uint8_t *p = b + 52 - 4 * i;
    p[0] = (uint8_t)(a[i] >> 24); p[1] = (uint8_t)(a[i] >> 16); p[2] = (uint8_t)(a[i] >> 8); p[3] = (uint8_t)a[i];
  }
}
/* x (n limbs, n <= 29) -> x mod r, by folding at bit 446: x = hi*2^446 + lo = hi*c + lo */
static void sc_reduce(uint32_t o[SL], const uint32_t *x, int n) {
  uint32_t cur[30], nxt[30];
  int len = n;
  memset(cur, 0, sizeof cur);
  memcpy(cur, x, (size_t)n * 4);
  while (len > SL || (len == SL && (cur[SL - 1] >> 30))) {
    /* hi = cur >> 446 ; lo = cur & (2^446 - 1) */
    uint32_t hi[17];
    int hl = len - 13; /* limbs of hi (bit 446 = limb 13 bit 30) */
    if (hl < 1) hl = 1;
    for (int i = 0; i < hl; i++) {
      uint32_t a = (13 + i < len) ? cur[13 + i] : 0, b = (14 + i < len) ? cur[14 + i] : 0;
      hi[i] = (a >> 30) | (b << 2);
    }
    memset(nxt, 0, sizeof nxt);
    for (int i = 0; i < 13; i++) nxt[i] = cur[i];
    nxt[13] = cur[13] & 0x3fffffffu;
    /* nxt += hi * c */
    for (int i = 0; i < hl; i++) {
      uint64_t cy = 0;
      for (int j = 0; j < 7; j++) {
        uint64_t t = (uint64_t)hi[i] * C_LIMBS[j] + nxt[i + j] + cy;
        nxt[i + j] = (uint32_t)t; cy = t >> 32;
      }
      for (int k = i + 7; cy; k++) { uint64_t t = (uint64_t)nxt[k] + cy; nxt[k] = (uint32_t)t; cy = t >> 32; }
    }
    memcpy(cur, nxt, sizeof cur);
    len = 30; while (len > 1 && cur[len - 1] == 0) len--;
  }
  /* cur < 2^446 < 2r : one conditional subtract */
  uint32_t t[SL]; uint64_t bw = 0;
  for (int i = 0; i < SL; i++) { uint64_t v = (uint64_t)cur[i] - R_LIMBS[i] - bw; t[i] = (uint32_t)v; bw = (v >> 63) & 1; }
  for (int i = 0; i < SL; i++) o[i] = bw ? cur[i] : t[i];
}
static void sc_mul_mod(uint32_t o[SL], const uint32_t a[SL], const uint32_t b[SL]) { /* Scalar::mul_mod / `*` */
  uint32_t p[2 * SL]; memset(p, 0, sizeof p);
  for (int i = 0; i < SL; i++) {
    uint64_t cy = 0;
    for (int j = 0; j < SL; j++) { uint64_t t = (uint64_t)a[i] * b[j] + p[i + j] + cy; p[i + j] = (uint32_t)t; cy = t >> 32; }
    p[i + SL] = (uint32_t)cy;
  }
  sc_reduce(o, p, 2 * SL);
}
static void sc_mul4_mod(uint32_t o[SL], const uint32_t a[SL]) { /* .mul_mod(Scalar::from(4)) */
  uint32_t four[SL] = {4}; sc_mul_mod(o, a, four);
}
static void sc_sub_mod(uint32_t o[SL], const uint32_t a[SL], const uint32_t b[SL]) { /* a, b in [0, r) */
  uint32_t t[SL]; uint64_t bw = 0;
  for (int i = 0; i < SL; i++) { uint64_t v = (uint64_t)a[i] - b[i] - bw; t[i] = (uint32_t)v; bw = (v >> 63) & 1; }
  uint64_t cy = 0;
  for (int i = 0; i < SL; i++) { uint64_t v = (uint64_t)t[i] + (bw ? R_LIMBS[i] : 0) + cy; o[i] = (uint32_t)v; cy = v >> 32; }
}

/* ====================================================================================
 * protocol: keygen / sign / verify / ecdh (src/ecc/ *.rs), one item at a time
 * ==================================================================================== */
static void gen_point(pt *g) { pt_from_affine_bytes(g, GEN_XY); }

/* s = 4 * BE(KMACXOF(pw,"",448,"SK",d)) mod r  (keypair.rs:42-43) -> 56 bytes BE */
static void secret_scalar_be(uint8_t s_be[56], const uint8_t *pw, size_t pwlen, int d) {
  uint8_t kb[56]; uint32_t a[SL], s[SL];
  capy_ref_kmac_xof(pw, pwlen, NULL, 0, 448, (const uint8_t *)"SK", 2, d, kb, 0);
  sc_from_be(a, kb); sc_mul4_mod(s, a); sc_to_be(s_be, s);
}

int capy_ref_ed448_scalar_mult(const uint8_t k_be[56], const uint8_t *pt_xy /* NULL = generator */, uint8_t out_xy[112]) {
  pt b, o;
  if (pt_xy) { int rc = pt_from_affine_bytes(&b, pt_xy); if (rc) return rc; } else gen_point(&b);
  pt_scalar_mult(&o, &b, k_be);
  pt_to_affine_bytes(out_xy, &o);
  return 0;
}
int capy_ref_ed448_keygen(const uint8_t *pw, size_t pwlen, int d, uint8_t out_xy[112]) { /* keypair.rs:41-51 */
  if (!valid_d(d)) return -1;
  uint8_t s_be[56];
  secret_scalar_be(s_be, pw, pwlen, d);
  return capy_ref_ed448_scalar_mult(s_be, NULL, out_xy);
}
int capy_ref_ed448_sign(const uint8_t *pw, size_t pwlen, const uint8_t *msg, size_t mlen, int d,
                        uint8_t h[56], uint8_t z_be[56]) { /* signable.rs:40-57 */
  if (!valid_d(d)) return -1;
  uint8_t s_be[56], kb[56], k_be[56], uxy[112];
  uint32_t s[SL], k[SL], a[SL], hs[SL], z[SL];
  secret_scalar_be(s_be, pw, pwlen, d);
  capy_ref_kmac_xof(s_be, 56, msg, mlen, 448, (const uint8_t *)"N", 1, d, kb, 0);
  sc_from_be(a, kb); sc_mul4_mod(k, a); sc_to_be(k_be, k);
  capy_ref_ed448_scalar_mult(k_be, NULL, uxy);
  capy_ref_kmac_xof(uxy, 56, msg, mlen, 448, (const uint8_t *)"T", 1, d, h, 0);
  sc_from_be(a, h); sc_from_be(s, s_be);
  sc_mul_mod(hs, a, s);
  sc_sub_mod(z, k, hs);
  sc_to_be(z_be, z);
  return 0;
}
int capy_ref_ed448_verify(const uint8_t pub_xy[112], const uint8_t *msg, size_t mlen, const uint8_t h[56],
                          const uint8_t z_be[56], int d) { /* signable.rs:72-86; 1 = ok, 0 = reject */
  if (!valid_d(d)) return -1;
  pt g, v, a, b, u;
  gen_point(&g);
  int rc = pt_from_affine_bytes(&v, pub_xy); if (rc) return rc;
  pt_scalar_mult(&a, &g, z_be);
  pt_scalar_mult(&b, &v, h);
  pt_add(&u, &a, &b);
  uint8_t uxy[112], hp[56];
  pt_to_affine_bytes(uxy, &u);
  capy_ref_kmac_xof(uxy, 56, msg, mlen, 448, (const uint8_t *)"T", 1, d, hp, 0);
  return memcmp(hp, h, 56) == 0;
}

/* ---- batches ---- */
int capy_ref_ed448_fixed_base_batch(const uint8_t *scalars_be56, uint64_t n, uint8_t *out_xy, int threads) {
  int t = clamp_threads(threads);
#pragma omp parallel for num_threads(t) schedule(dynamic, 4) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++) capy_ref_ed448_scalar_mult(scalars_be56 + 56 * i, NULL, out_xy + 112 * i);
  return 0;
}
int capy_ref_ed448_var_base_batch(const uint8_t *scalars_be56, const uint8_t *pts_xy, uint64_t n, uint8_t *out_xy,
                                  int threads) {
  int t = clamp_threads(threads), bad = 0;
#pragma omp parallel for num_threads(t) schedule(dynamic, 4) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++)
    if (capy_ref_ed448_scalar_mult(scalars_be56 + 56 * i, pts_xy + 112 * i, out_xy + 112 * i)) {
#pragma omp atomic write
      bad = 1;
    }
  return bad ? -4 : 0;
}
int capy_ref_ed448_keygen_batch(const uint8_t *pws, const uint64_t *pw_off, uint64_t n, int d, uint8_t *out_xy, int threads) {
  if (!valid_d(d)) return -1;
  int t = clamp_threads(threads);
#pragma omp parallel for num_threads(t) schedule(dynamic, 4) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++)
    capy_ref_ed448_keygen(pws + pw_off[i], (size_t)(pw_off[i + 1] - pw_off[i]), d, out_xy + 112 * i);
  return 0;
}
int capy_ref_ed448_sign_batch(const uint8_t *pws, const uint64_t *pw_off, const uint8_t *msgs, const uint64_t *msg_off,
                              uint64_t n, int d, uint8_t *h56, uint8_t *z56, int threads) {
  if (!valid_d(d)) return -1;
  int t = clamp_threads(threads);
#pragma omp parallel for num_threads(t) schedule(dynamic, 4) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++)
    capy_ref_ed448_sign(pws + pw_off[i], (size_t)(pw_off[i + 1] - pw_off[i]), msgs + msg_off[i],
                        (size_t)(msg_off[i + 1] - msg_off[i]), d, h56 + 56 * i, z56 + 56 * i);
  return 0;
}
int capy_ref_ed448_verify_batch(const uint8_t *pub_xy, const uint8_t *msgs, const uint64_t *msg_off, const uint8_t *h56,
                                const uint8_t *z56, uint64_t n, int d, uint8_t *ok, int threads) {
  if (!valid_d(d)) return -1;
  int t = clamp_threads(threads);
#pragma omp parallel for num_threads(t) schedule(dynamic, 4) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    int r = capy_ref_ed448_verify(pub_xy + 112 * i, msgs + msg_off[i], (size_t)(msg_off[i + 1] - msg_off[i]),
                                  h56 + 56 * i, z56 + 56 * i, d);
    ok[i] = (uint8_t)(r == 1);
  }
  return 0;
}
/* ECDH core (encryptable.rs:36-38,76-78): W.x for W = [4*BE(k56) mod r] * V ; out 56 bytes LE */
int capy_ref_ed448_ecdh_batch(const uint8_t *k_rand56, const uint8_t *pub_xy, uint64_t n, uint8_t *wx56, int threads) {
  int t = clamp_threads(threads), bad = 0;
#pragma omp parallel for num_threads(t) schedule(dynamic, 4) if (t > 1)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    uint32_t a[SL], k[SL]; uint8_t k_be[56], xy[112];
    sc_from_be(a, k_rand56 + 56 * i); sc_mul4_mod(k, a); sc_to_be(k_be, k);
    if (capy_ref_ed448_scalar_mult(k_be, pub_xy + 112 * i, xy)) {
#pragma omp atomic write
      bad = 1;
    }
    memcpy(wx56 + 56 * i, xy, 56);
  }
  return bad ? -4 : 0;
}

/* scalar helpers exposed for unit tests of the mod-r arithmetic */
void capy_ref_sc_mul_mod(const uint8_t a_be[56], const uint8_t b_be[56], uint8_t o_be[56]) {
  uint32_t a[SL], b[SL], o[SL]; sc_from_be(a, a_be); sc_from_be(b, b_be); sc_mul_mod(o, a, b); sc_to_be(o_be, o);
}
void capy_ref_fe_mul(const uint8_t a[56], const uint8_t b[56], uint8_t o[56]) {
  fe x, y, z; fe_from_bytes(&x, a); fe_from_bytes(&y, b); fe_mul(&z, &x, &y); fe_to_bytes(o, &z);
}
void capy_ref_fe_inv(const uint8_t a[56], uint8_t o[56]) {
  fe x, z; fe_from_bytes(&x, a); fe_inv(&z, &x); fe_to_bytes(o, &z);
}
