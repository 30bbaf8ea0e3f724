/*
 * capy_gpu.h -- C ABI of libcapycrypt_gpu: the B200 (sm_100a) batch engine for capyCRYPT's two
 * data-parallel hot paths (Keccak sponge -> SHA3/cSHAKE/KMACXOF; Ed448-Goldilocks scalar
 * multiplication -> keygen / Schnorr sign+verify / ECDH core).
 *
 * The reference (capyCRYPT 0.7.5, pure Rust) has NO plugin or FFI boundary: its operator
 * surface is a set of traits on `Message` plus `KeyPair::new` and `pub fn kmac_xof`.  Every
 * entry point below is the batched form of one of those operators and cites the reference
 * interface it replaces (paths relative to the reference tree).  A `capycrypt::gpu` Rust
 * module binds these with `extern "C"` (see INTEGRATION.md and rust/gpu.rs).
 *
 * Conventions
 *  - plain pointers and sizes only; every call returns an int status (CAPY_OK == 0, errors < 0);
 *    nothing throws or aborts across the boundary.
 *  - the caller owns every buffer.  Host entry points are blocking; host pointers may be
 *    pageable: large pageable buffers are staged through page-locked buffers of the ctx by a few
 *    copy threads (env CAPY_COPY_THREADS, default up to 8) -- about 20 GB/s; memory from
 *    capy_host_alloc goes straight to the DMA engines -- about 47 GB/s.  `_dev` twins take DEVICE
 *    pointers plus a CUDA stream and are asynchronous on that stream.
 *  - batches: `data` is a packed byte array and `off` has n+1 uint64 offsets into it (item i
 *    = data[off[i] .. off[i+1]) ).  `_fixed` variants take one length and a stride instead.
 *  - d_bits is the reference's SecParam (lib.rs:113-135): 224, 256, 384 or 512; anything else
 *    returns CAPY_ERR_BAD_SECPARAM (= OperationError::UnsupportedSecurityParameter).
 *  - results are bit-exact with the reference INCLUDING its documented deviations from
 *    FIPS 202 / SP 800-185.
 *  - field elements: 56 bytes little-endian canonical (FieldElement::to_bytes);
 *    scalars: 56 bytes big-endian (aux_functions.rs:102-110); points: affine x || y, 112 bytes.
 *  - deterministic: no RNG inside; nonces are inputs.
 *  - `_dev` calls only enqueue work (ragged batches also read one small summary back to plan their launch).
 *    They share the ctx's device scratch buffers, so the `_dev` calls of one ctx must be issued on ONE stream at a
 *    time; use one ctx per concurrently used stream.
 *  - a ctx may span several GPUs; host entry points shard the batch across them by contiguous
 *    ranges (no collective: items are independent).  Calls on one ctx are serialised by an
 *    internal mutex; distinct ctxs are independent.
 */
#ifndef CAPY_GPU_H
#define CAPY_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPY_OK 0
#define CAPY_ERR_BAD_SECPARAM (-1) /* lib.rs:126-134 UnsupportedSecurityParameter */
#define CAPY_ERR_BAD_ARG (-2)
#define CAPY_ERR_CUDA (-3)
#define CAPY_ERR_BAD_POINT (-4) /* an input point is not on the curve */
#define CAPY_ERR_NO_DEVICE (-5)
#define CAPY_ERR_OOM (-6)

/* flags */
#define CAPY_FLAG_NONE 0u
/* ragged batches are processed longest-message-first (device-side counting sort + one tiny D2H sync);
 * this flag keeps the caller's order and makes the _dev call fully asynchronous */
#define CAPY_FLAG_NO_SORT 1u
/* chain-bound ragged batches run their longest messages with two threads per message (the chain of one message
 * advances faster); this flag keeps one thread per message everywhere (A/B measurements) */
#define CAPY_FLAG_NO_PAIR 2u
/* diagnostics: force the number of warp-tier chains that share a warp scheduler (1..3) in a chain-bound ragged batch
 * instead of letting the planner choose: flags |= c << CAPY_FLAG_WARP_COSCHED_SHIFT */
#define CAPY_FLAG_WARP_COSCHED_SHIFT 8

typedef struct capy_ctx capy_ctx;

/* ---- context ---------------------------------------------------------------------------- */
/* devices == NULL && n_devices <= 0: use the current CUDA device only. */
int capy_gpu_init(const int* devices, int n_devices, capy_ctx** out_ctx);
void capy_gpu_destroy(capy_ctx* ctx);
int capy_gpu_device_count(const capy_ctx* ctx);
/* Overwrites every device scratch buffer of the ctx -- and the page-locked host buffers pageable inputs were staged
 * through -- with zeros.  The pipelines keep intermediate secrets there (derived scalars, KMAC key material, staged
 * passwords) until the next call reuses the buffer; the reference does not zeroize either, so this is an extra for
 * callers who want it.  Blocking. */
int capy_gpu_scrub(capy_ctx* ctx);
const char* capy_strerror(int status);
/* last CUDA error text seen by this ctx (for CAPY_ERR_CUDA) */
const char* capy_last_cuda_error(const capy_ctx* ctx);
int capy_version(void);
/* pinned host memory helpers (optional; faster H2D/D2H for the host entry points) */
void* capy_host_alloc(size_t bytes);
void capy_host_free(void* p);
/* Measurement aid: the ceiling of the host <-> device link for this process.  Per repetition one plain cudaMemcpyAsync
 * of in_bytes host -> device and one of out_bytes device -> host, on two streams so the two directions overlap, no
 * kernel; ms_per_rep is device-timed (events).  Buffers from capy_host_alloc give the pinned-memory figure.  bench.py
 * reports the engine's end-to-end throughput as a fraction of this. */
int capy_copy_probe(capy_ctx* ctx, int dev_index, const void* h_in, size_t in_bytes, void* h_out, size_t out_bytes,
                    int reps, double* ms_per_rep);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches evidence) */
uint64_t capy_launch_count(const capy_ctx* ctx);
/* Plan cache for ragged batches.  A ragged sponge call orders its items longest-first and, for chain-bound batches,
 * splits them into tiers; the shape of that launch depends on the length histogram, which the host reads back (two
 * small D2H copies + stream synchronisations).  With the cache enabled the plan of an offsets array is kept, keyed by
 * (device pointer, n, unit), and every later call that passes the same array -- _dev entry points: the caller's device
 * array; cSHAKE / KMAC / AE alike -- launches without touching the host: the call is asynchronous on the caller's
 * stream.  Contract: an offsets array passed again under the same address has not changed.  A stale plan costs speed,
 * never correctness (the work order stays a permutation of the items and any item may run in any tier).  At most 16
 * plans per device (least recently used evicted); enable = 0 drops them.  Off by default. */
int capy_gpu_set_plan_cache(capy_ctx* ctx, int enable);

/* ---- diagnostics: the host-side launch planner of chain-bound ragged sponge batches (no GPU needed) ------- */
/* items_longer_than[k] = number of items whose message has MORE than k whole rate blocks (k < n_bins), n items,
 * the longest with max_blocks blocks, total_blocks in all.  Returns how many of the longest items run with a whole
 * warp per item and how many of the next with two threads per item (DESIGN.md "chain-bound batches"). */
int capy_plan_tiers(const uint32_t* items_longer_than, uint32_t n_bins, uint64_t n, uint32_t max_blocks,
                    uint64_t total_blocks, int sm_count, uint64_t* warp_items, uint64_t* pair_items);
/* same, with the warp-per-item count split by how many such chains share a warp scheduler: warp_items_by_sharing[c - 1]
 * items run c chains per scheduler, c = 1..3 (a warp-tier chain uses a fraction of its scheduler's issue slots; the
 * longest chains get a scheduler to themselves, the next ones share) */
int capy_plan_tiers3(const uint32_t* items_longer_than, uint32_t n_bins, uint64_t n, uint32_t max_blocks,
                     uint64_t total_blocks, int sm_count, uint64_t* warp_items_by_sharing, uint64_t* pair_items);

/* ---- diagnostics: chain cutting of uniform cSHAKE / KMAC batches (no GPU needed) ------------------------------------
 * A uniform batch whose warps fill the schedulers unevenly (2^16 items on 148 SMs = 3.46 warps per scheduler) is run
 * as two dependent jobs in one launch: the first absorbs `cut_after_blocks` blocks of every item and hands the state
 * over, the second finishes (twice the warps, half as long each).  Returns 1 and the cut when the engine would do that
 * for n items that absorb absorb_blocks blocks (after the cached prefix) and run squeeze_extra permutations between
 * squeeze blocks, 0 when the batch is launched as it is. */
int capy_chain_cut(int sm_count, uint64_t n, uint64_t absorb_blocks, uint64_t squeeze_extra, uint64_t* cut_after_blocks);
/* (environment, read at every launch: CAPY_NO_CHAIN_SPLIT=1 launches such batches uncut -- the A/B switch of the probes
 * and of tests/test_chain_split_gpu.py; results are the same bytes either way) */

/* ---- diagnostics: how a ragged sponge batch is divided over the devices of a ctx (no GPU needed) ------------------- */
/* owner[i] = index of the device that hashes item i when capy_sha3_batch runs on a ctx of `parts` devices: the outliers
 * (messages that cost >= 8 x the average, cost = len / unit_bytes + per_item_cost permutations) are dealt out longest
 * first to the least loaded device (LPT over the chains), the rest follows in contiguous index ranges that
 * top every device up to the same load. */
int capy_lpt_shares(const uint64_t* off, uint64_t n, uint32_t parts, uint32_t unit_bytes, uint64_t per_item_cost,
                    uint32_t* owner);

/* ---- SHA3-d : SpongeHashable::compute_sha3_hash (sha3/hashable.rs:19-21 -> shake,
 *      sha3/shake_functions.rs:24-32 -> sponge_absorb/squeeze, sha3/sponge.rs:10-34) ------- */
/* digests: n * (d_bits/8) bytes, item-major.  Hashes the ORIGINAL bytes of every message (the
 * reference additionally leaves Message.msg suffixed+padded, quirk Q5; the Rust shim may
 * replicate that append host-side). */
int capy_sha3_batch(capy_ctx* ctx, int d_bits, const uint8_t* data, const uint64_t* off, uint64_t n,
                    uint8_t* digests, uint32_t flags);
int capy_sha3_batch_fixed(capy_ctx* ctx, int d_bits, const uint8_t* data, uint64_t msg_len, uint64_t stride,
                          uint64_t n, uint8_t* digests, uint32_t flags);
int capy_sha3_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_data,
                        const uint64_t* d_off, uint64_t n, uint8_t* d_digests, uint32_t flags);
int capy_sha3_batch_fixed_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_data,
                              uint64_t msg_len, uint64_t stride, uint64_t n, uint8_t* d_digests, uint32_t flags);

/* ---- cSHAKE : cshake (sha3/shake_functions.rs:49-64), capacity = d bits (quirk Q7) ---------- */
/* out: n * (out_bits/8) bytes.  N = S = "" reproduces the reference's quirk Q4. */
int capy_cshake_batch(capy_ctx* ctx, int d_bits, const uint8_t* data, const uint64_t* off, uint64_t n,
                      const uint8_t* fn_name, uint32_t fn_len, const uint8_t* custom, uint32_t custom_len,
                      uint64_t out_bits, uint8_t* out);
int capy_cshake_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_data,
                          const uint64_t* d_off, uint64_t n, const uint8_t* fn_name, uint32_t fn_len,
                          const uint8_t* custom, uint32_t custom_len, uint64_t out_bits, uint8_t* d_out);

/* ---- KMACXOF : pub fn kmac_xof (sha3/shake_functions.rs:79-89); also
 *      SpongeHashable::compute_tagged_hash (sha3/hashable.rs:33-35) with out_bits = d_bits ---- */
/* keys/key_off: per-item keys (packed + n+1 offsets).  out_off == NULL: every item gets
 * out_bits/8 bytes at stride out_bits/8; otherwise item i writes out[out_off[i]..out_off[i+1])
 * (variable-length squeeze, e.g. the keystream shape of sha3/encryptable.rs:41). */
int capy_kmac_xof_batch(capy_ctx* ctx, int d_bits, const uint8_t* keys, const uint64_t* key_off,
                        const uint8_t* data, const uint64_t* off, uint64_t n, const uint8_t* custom,
                        uint32_t custom_len, uint64_t out_bits, const uint64_t* out_off, uint8_t* out);
int capy_kmac_xof_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_keys,
                            const uint64_t* d_key_off, const uint8_t* d_data, const uint64_t* d_off, uint64_t n,
                            const uint8_t* custom, uint32_t custom_len, uint64_t out_bits,
                            const uint64_t* d_out_off, uint8_t* d_out);
/* fixed-size twin: every key is key_len bytes at key_stride, every message msg_len at msg_stride */
int capy_kmac_xof_batch_fixed_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_keys,
                                  uint64_t key_len, uint64_t key_stride, const uint8_t* d_data, uint64_t msg_len,
                                  uint64_t msg_stride, uint64_t n, const uint8_t* custom, uint32_t custom_len,
                                  uint64_t out_bits, uint8_t* d_out);

/* ---- FIPS 202 SHAKE128/256 -- NO reference counterpart (the reference's `shake` is SHA3-d);
 *      provided because BASELINE.json config 2 names SHAKE256; checked against hashlib. ------ */
int capy_fips_shake_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int shake_bits /*128|256*/,
                              const uint8_t* d_data, const uint64_t* d_off, uint64_t n, uint64_t out_bytes,
                              uint8_t* d_out);

/* ---- Ed448 scalar multiplication : `ExtendedPoint * Scalar` of tiny_ed448_goldilocks 0.1.8,
 *      call sites ecc/keypair.rs:44, ecc/signable.rs:48,77, ecc/encryptable.rs:37-38,78 ------- */
/* out[i] = [k_i]G for the exact integer k_i = BE(scalars[56*i..]) (reduced mod r internally,
 * G has order r).  Fixed-base comb, constant-time table scan. */
int capy_ed448_fixed_base_batch(capy_ctx* ctx, const uint8_t* scalars_be56, uint64_t n, uint8_t* out_xy112);
int capy_ed448_fixed_base_batch_dev(capy_ctx* ctx, int dev_index, void* stream, const uint8_t* d_scalars_be56,
                                    uint64_t n, uint8_t* d_out_xy112);
/* out[i] = [k_i]P_i with k_i the exact (unreduced, up to 448-bit) integer (quirk Q10).
 * Returns CAPY_ERR_BAD_POINT if any P_i is off-curve (that item's output is all zero). */
int capy_ed448_var_base_batch(capy_ctx* ctx, const uint8_t* scalars_be56, const uint8_t* points_xy112, uint64_t n,
                              uint8_t* out_xy112);
int capy_ed448_var_base_batch_dev(capy_ctx* ctx, int dev_index, void* stream, const uint8_t* d_scalars_be56,
                                  const uint8_t* d_points_xy112, uint64_t n, uint8_t* d_out_xy112,
                                  int* d_bad_flag /* device int, set non-zero on a bad point; may be NULL */);

/* ---- KeyPair::new public-key derivation (ecc/keypair.rs:41-51) ------------------------------- */
/* s = 4 * BE(KMACXOF(pw,"",448,"SK",d)) mod r ; V = [s]G ; out = affine V. */
int capy_ed448_keygen_batch(capy_ctx* ctx, int d_bits, const uint8_t* pws, const uint64_t* pw_off, uint64_t n,
                            uint8_t* out_xy112);
int capy_ed448_keygen_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_pws,
                                const uint64_t* d_pw_off, uint64_t n, uint8_t* d_out_xy112);

/* ---- Signable::sign / verify (ecc/signable.rs:40-57, 72-86) ------------------------------------ */
/* sign: h56[i] (56 bytes) and z_be56[i] (56 bytes big-endian, canonical in [0, r)). */
int capy_ed448_sign_batch(capy_ctx* ctx, int d_bits, const uint8_t* pws, const uint64_t* pw_off,
                          const uint8_t* msgs, const uint64_t* msg_off, uint64_t n, uint8_t* h56, uint8_t* z_be56);
int capy_ed448_sign_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_pws,
                              const uint64_t* d_pw_off, const uint8_t* d_msgs, const uint64_t* d_msg_off,
                              uint64_t n, uint8_t* d_h56, uint8_t* d_z_be56);
/* verify: ok[i] = 1 iff the signature verifies (Ok(())), 0 = SignatureVerificationFailure.
 * Off-curve public keys give ok = 0 and the call returns CAPY_ERR_BAD_POINT. */
int capy_ed448_verify_batch(capy_ctx* ctx, int d_bits, const uint8_t* pub_xy112, const uint8_t* msgs,
                            const uint64_t* msg_off, const uint8_t* h56, const uint8_t* z_be56, uint64_t n,
                            uint8_t* ok);
int capy_ed448_verify_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_pub_xy112,
                                const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_h56,
                                const uint8_t* d_z_be56, uint64_t n, uint8_t* d_ok, int* d_bad_flag);

/* ---- ECDH core of KeyEncryptable (ecc/encryptable.rs:36-38 and :76-78) -------------------------- */
/* k_i = 4 * BE(k_rand56[i]) mod r ; wx56[i] = x([k_i]V_i) ; if z_xy112 != NULL also Z_i = [k_i]G.
 * Decrypt side: pass the nonce points Z_i as `pub` and the secret scalars via capy_ed448_var_base_batch. */
int capy_ed448_ecdh_batch(capy_ctx* ctx, const uint8_t* k_rand56, const uint8_t* pub_xy112, uint64_t n,
                          uint8_t* wx56, uint8_t* z_xy112);

/* ---- Sponge AE : SpongeEncryptable::sha3_encrypt / sha3_decrypt (sha3/encryptable.rs:29-45, 58-83) and the
 *      symmetric half of KEMEncryptable::kem_encrypt / kem_decrypt (kem/encryptable.rs:51-57, 91-104, where
 *      `pw` is the ML-KEM shared secret; ML-KEM itself is out of scope) --------------------------------- */
#define CAPY_AE_SHA3 0 /* customisation strings "S", "SKE", "SKA" */
#define CAPY_AE_KEM 1  /* "S", "KEMKE", "KEMKA" */
/* (ke || ka) = KMACXOF(z_i || pw_i, "", 1024, "S"); tag64[i] = KMACXOF(ka, m_i, 512, KA);
 * ct_i = KMACXOF(ke, "", |m_i|, KE) xor m_i, written at the message's own offsets (|ct| == |m|).
 * nonces: n * nonce_len bytes (the reference draws nonce_len = 512 random BYTES, :31); no RNG inside. */
int capy_sponge_encrypt_batch(capy_ctx* ctx, int d_bits, int variant, const uint8_t* pws, const uint64_t* pw_off,
                              const uint8_t* nonces, uint64_t nonce_len, const uint8_t* msgs, const uint64_t* msg_off,
                              uint64_t n, uint8_t* ct, uint8_t* tag64);
/* ok[i] = 1 for Ok(()); 0 = OperationError::SHA3DecryptionFailure, and out_i is then the ciphertext again
 * (the reference XORs the keystream back, :77-82).  `out` must not alias `ct`. */
int capy_sponge_decrypt_batch(capy_ctx* ctx, int d_bits, int variant, const uint8_t* pws, const uint64_t* pw_off,
                              const uint8_t* nonces, uint64_t nonce_len, const uint8_t* ct, const uint64_t* ct_off,
                              const uint8_t* tag64, uint64_t n, uint8_t* out, uint8_t* ok);
/* device twins; pw_bytes = total bytes in d_pws (sizes the z || pw key buffer) */
int capy_sponge_encrypt_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, int variant, const uint8_t* d_pws,
                                  const uint64_t* d_pw_off, uint64_t pw_bytes, const uint8_t* d_nonces, uint64_t nonce_len,
                                  const uint8_t* d_msgs, const uint64_t* d_msg_off, uint64_t n, uint8_t* d_ct,
                                  uint8_t* d_tag64);
int capy_sponge_decrypt_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, int variant, const uint8_t* d_pws,
                                  const uint64_t* d_pw_off, uint64_t pw_bytes, const uint8_t* d_nonces, uint64_t nonce_len,
                                  const uint8_t* d_ct, const uint64_t* d_ct_off, const uint8_t* d_tag64, uint64_t n,
                                  uint8_t* d_out, uint8_t* d_ok);

/* ---- ECDHIES : KeyEncryptable::key_encrypt / key_decrypt (ecc/encryptable.rs:34-50, 72-94) ---------------- */
/* k_i = 4 * BE(k_rand56[i]) mod r; W = [k]V_i; z_xy112[i] = [k]G (Message.asym_nonce, affine);
 * (ke || ka) = KMACXOF(W.x, "", 896, "PK"); tag56[i] = KMACXOF(ka, m_i, 448, "PKA");
 * ct_i = KMACXOF(ke, "", |m_i|, "PKE") xor m_i.  Off-curve V_i -> CAPY_ERR_BAD_POINT (W.x taken as 0). */
int capy_ed448_key_encrypt_batch(capy_ctx* ctx, int d_bits, const uint8_t* pub_xy112, const uint8_t* k_rand56,
                                 const uint8_t* msgs, const uint64_t* msg_off, uint64_t n, uint8_t* ct, uint8_t* tag56,
                                 uint8_t* z_xy112);
/* s = 4 * BE(KMACXOF(pw, "", 448, "SK")) mod r; W = [s]Z_i; ok[i] = 1 for Ok(()), 0 = KeyDecryptionError with
 * out_i = ciphertext (:88-93).  `out` must not alias `ct`. */
int capy_ed448_key_decrypt_batch(capy_ctx* ctx, int d_bits, const uint8_t* pws, const uint64_t* pw_off,
                                 const uint8_t* z_xy112, const uint8_t* ct, const uint64_t* ct_off, const uint8_t* tag56,
                                 uint64_t n, uint8_t* out, uint8_t* ok);
int capy_ed448_key_encrypt_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_pub_xy112,
                                     const uint8_t* d_k_rand56, const uint8_t* d_msgs, const uint64_t* d_msg_off, uint64_t n,
                                     uint8_t* d_ct, uint8_t* d_tag56, uint8_t* d_z_xy112, int* d_bad_flag);
int capy_ed448_key_decrypt_batch_dev(capy_ctx* ctx, int dev_index, void* stream, int d_bits, const uint8_t* d_pws,
                                     const uint64_t* d_pw_off, const uint8_t* d_z_xy112, const uint8_t* d_ct,
                                     const uint64_t* d_ct_off, const uint8_t* d_tag56, uint64_t n, uint8_t* d_out,
                                     uint8_t* d_ok, int* d_bad_flag);

#ifdef __cplusplus
}
#endif
#endif /* CAPY_GPU_H */
