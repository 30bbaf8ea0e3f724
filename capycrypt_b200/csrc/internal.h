// internal.h -- host-side context shared by the translation units of libcapycrypt_gpu.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/capy_gpu.h"

namespace capy {

constexpr int kNumStreams = 3;
constexpr int kNumScratch = 80;  // 0..23 sponge entry points, 24..55 Ed448 pipelines, 56..79 AE pipelines

// grow-only device scratch slots; each API call uses a fixed set of slot ids
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
};

struct PrefixState {
  uint64_t* d_state = nullptr;   // 25 lanes after the whole-block part of the prefix
  uint8_t* d_prefix = nullptr;   // prefix bytes on the device
  uint32_t prefix_len = 0;
  uint32_t skip_blocks = 0;
};

struct Ed448Tables;  // defined in ed448_api.cu

struct DeviceCtx {
  int dev = 0;
  cudaStream_t streams[kNumStreams] = {};
  Scratch scratch[kNumScratch];
  std::map<std::string, PrefixState> prefix_cache;
  Ed448Tables* ed = nullptr;
  int sm_count = 0;
};

}  // namespace capy

struct capy_ctx {
  std::vector<capy::DeviceCtx> devs;
  std::mutex mu;
  std::string last_cuda_error;
  std::atomic<uint64_t> launches{0};
};

namespace capy {

// RAII device switch
struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

int cuda_fail(capy_ctx* ctx, cudaError_t e, const char* what);
#define CAPY_CUDA(ctx, expr)                                             \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return ::capy::cuda_fail((ctx), _e, #expr);   \
  } while (0)

// returns nullptr on OOM
void* scratch_get(DeviceCtx& dc, int slot, size_t bytes);

inline bool valid_secparam(int d) { return d == 224 || d == 256 || d == 384 || d == 512; }
inline uint32_t bytepad_value(int d) { return d == 224 ? 172u : d == 256 ? 168u : d == 384 ? 152u : 136u; }  // lib.rs:137-144
inline uint32_t sha3_capacity(int d) {  // constants.rs:38-45
  int x = d * 2;
  return x <= 448 ? 448u : x <= 512 ? 512u : x <= 768 ? 768u : 1024u;
}

// ---- device-level launchers (async on `stream`), implemented in sha3_api.cu -----------------
struct KmacDevArgs {
  int d_bits;
  const uint8_t* keys;
  const uint64_t* key_off;  // nullptr -> fixed
  uint64_t key_len, key_stride;
  const uint8_t* data;
  const uint64_t* off;  // nullptr -> fixed
  uint64_t msg_len, msg_stride;
  uint64_t n;
  const uint8_t* custom;
  uint32_t custom_len;
  uint64_t out_bytes;
  const uint64_t* out_off;
  uint64_t out_stride;
  uint8_t* out;
  const uint8_t* xor_in = nullptr;  // out = keystream ^ xor_in (same layout as out)
  bool no_sort = false;  // keep the caller's order (skips the length bucketing and its tiny D2H sync)
};
int launch_kmac_xof(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& a);

}  // namespace capy
