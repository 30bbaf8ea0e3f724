// internal.h -- host-side context shared by the translation units of libcapycrypt_gpu.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/capy_gpu.h"

namespace capy {

constexpr int kNumStreams = 3;
constexpr int kNumScratch = 152;  // 0..23 sponge entry points, 24..55 Ed448 pipelines, 56..79 AE pipelines, 80..143 Ed448 slots of streams 1, 2,
                                  // 144..151 chain hand-off (states, tickets) of the sponge entry points

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
  uint64_t last_use = 0;         // LRU stamp (the customisation string is caller-controlled: the cache is bounded)
};
constexpr size_t kMaxPrefixEntries = 64;

// How a ragged batch is launched (sha3_api.cu: plan_ragged).  `order` = work order (longest first) or nullptr.
struct LaunchPlan {
  const uint32_t* order = nullptr;
  int warps_per_smsp = 0;   // 0 = unthrottled
  uint64_t warp_items[3] = {0, 0, 0};  // one warp per item: [c - 1] of them run c chains per scheduler (ranks in this order)
  uint64_t pair_items = 0;             // the next ranks: two threads per item
  uint64_t uniform_blocks = 0;         // != 0: every item holds this many whole blocks (+ a partial one): nothing to order
};

// Cached plan of a ragged batch, keyed by the caller's offsets array (capy_gpu_set_plan_cache): a repeated call with the
// same (device pointer, n, unit) launches without touching the host.  A stale plan (the caller rewrote the offsets in
// place) costs speed only: `order` stays a permutation of [0, n) and any item may run in any tier.
struct PlanEntry {
  const uint64_t* off = nullptr;
  uint64_t n = 0;
  uint32_t unit = 0;
  bool allow_pair = true;
  LaunchPlan plan;
  uint32_t* owned_order = nullptr;
  uint64_t last_use = 0;
};
constexpr size_t kMaxPlanEntries = 16;

// persistent worker thread of one device of a multi-device ctx: host entry points post their per-device closure here
// instead of spawning a std::thread per call
class DeviceWorker {
 public:
  explicit DeviceWorker(int dev);
  ~DeviceWorker();
  void post(std::function<void()> f);

 private:
  void loop(int dev);
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  bool stop_ = false;
  std::thread th_;
};

struct Ed448Tables;  // defined in ed448_api.cu

// ---- staging of pageable host memory ---------------------------------------------------------------------------------
// A copy between the device and ordinary (pageable) host memory goes through the driver's own bounce buffers on the
// calling thread: 8.9 GB/s for the 64 MiB + 32 MiB of BASELINE config 1 against 46.7 GB/s from page-locked buffers.  Host
// entry points therefore stage large pageable buffers themselves: a few host threads copy a piece into (out of) a
// page-locked double buffer while the DMA of the previous piece runs.  Buffers from capy_host_alloc skip all of this.
constexpr size_t kStagePiece = 8ull << 20;      // bytes per staging half
constexpr size_t kStageMinBytes = 256ull << 10;  // smaller copies are left to the driver
struct StageSlot {
  uint8_t* p = nullptr;  // 2 x kStagePiece, page-locked
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool pending[2] = {false, false};
  // device-to-host: where the bytes of a half go once its copy has arrived
  void* drain_dst[2] = {nullptr, nullptr};
  size_t drain_bytes[2] = {0, 0};
  int next = 0;
};

// a few persistent helper threads that split one memcpy between them (and the caller)
class CopyPool {
 public:
  explicit CopyPool(int helpers);
  ~CopyPool();
  void copy(void* dst, const void* src, size_t bytes);

 private:
  void loop();
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> th_;
  uint8_t* dst_ = nullptr;
  const uint8_t* src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  int next_ = 0, slices_ = 0, left_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

struct DeviceCtx {
  int dev = 0;
  int index = 0;  // position in capy_ctx::devs
  cudaStream_t streams[kNumStreams] = {};
  Scratch scratch[kNumScratch];
  std::map<std::string, PrefixState> prefix_cache;
  std::vector<PlanEntry> plans;
  // host copies of the offsets arrays a host entry point has staged into device scratch (stage_packed), keyed by the
  // device address: plan_ragged reads the lengths from the host copy instead of fetching the histogram summary from the
  // device (no stream synchronisation inside a host-buffer call).  Valid only while that host call runs (HostOffScope).
  struct HostOff {
    const uint64_t* d_off;
    const uint64_t* h_off;
    uint64_t count;  // entries (items + 1)
  };
  std::vector<HostOff> host_offs;
  uint64_t use_clock = 0;
  Ed448Tables* ed = nullptr;
  int sm_count = 0;
  // host-side state of THIS device (scratch slots, caches, tables): one lock per device, so callers that use
  // different devices of a ctx do not serialise
  std::unique_ptr<std::mutex> mu{new std::mutex()};
  std::unique_ptr<DeviceWorker> worker;  // only in a multi-device ctx
  // staging of pageable host buffers, one double buffer per scratch slot that needs it (created at first use)
  std::vector<StageSlot> stage{(size_t)kNumScratch};
  std::unique_ptr<CopyPool> copy_pool;
};

}  // namespace capy

struct capy_ctx {
  std::vector<capy::DeviceCtx> devs;
  std::mutex err_mu;  // guards last_cuda_error only
  std::string last_cuda_error;
  std::atomic<uint64_t> launches{0};
  std::atomic<int> plan_cache{0};
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

// host <-> device copies of the host entry points (hostbatch.h); `slot` = the scratch slot of the device buffer
int copy_in(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t st, int slot, void* d_dst, const void* h_src, size_t bytes);
int copy_out(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t st, int slot, void* h_dst, const void* d_src, size_t bytes);
// waits for the staged device-to-host copies of this device and hands their bytes to the caller's buffers
int stage_drain(capy_ctx* ctx, DeviceCtx& dc);
void stage_free(DeviceCtx& dc);

// the host arrays registered in dc.host_offs belong to the caller of ONE host entry point: forget them when it returns
struct HostOffScope {
  DeviceCtx& dc;
  explicit HostOffScope(DeviceCtx& d) : dc(d) { dc.host_offs.clear(); }
  ~HostOffScope() { dc.host_offs.clear(); }
};
inline void host_off_register(DeviceCtx& dc, const uint64_t* d_off, const uint64_t* h_off, uint64_t count) {
  for (auto& e : dc.host_offs)
    if (e.d_off == d_off) {
      e.h_off = h_off;
      e.count = count;
      return;
    }
  dc.host_offs.push_back({d_off, h_off, count});
}
inline const uint64_t* host_off_find(const DeviceCtx& dc, const uint64_t* d_off, uint64_t n) {
  for (const auto& e : dc.host_offs)
    if (e.d_off == d_off && e.count >= n + 1) return e.h_off;
  return nullptr;
}

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
// two passes over the same items (same d, same n) as one launch; the plan (work order) is taken from `a`
int launch_kmac_xof2(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& a, const KmacDevArgs& b);
// two passes where `second` absorbs what `first` wrote for the same item: one launch of dependent jobs when the batch is large
int launch_kmac_xof_dep(capy_ctx* ctx, DeviceCtx& dc, cudaStream_t stream, const KmacDevArgs& first, const KmacDevArgs& second);

}  // namespace capy
