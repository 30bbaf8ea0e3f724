// ctx.cu -- context, streams, scratch memory and error plumbing of libcapycrypt_gpu.
#include <cstdio>
#include <cstring>

#include "internal.h"

namespace capy {

int cuda_fail(capy_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) {
    ctx->last_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + what;
  }
  cudaGetLastError();  // clear sticky-less errors
  return e == cudaErrorMemoryAllocation ? CAPY_ERR_OOM : CAPY_ERR_CUDA;
}

void* scratch_get(DeviceCtx& dc, int slot, size_t bytes) {
  Scratch& s = dc.scratch[slot];
  if (bytes == 0) bytes = 16;
  if (s.cap >= bytes) return s.p;
  if (s.p) {
    // other streams may still read the old buffer
    cudaDeviceSynchronize();
    cudaFree(s.p);
    s.p = nullptr;
    s.cap = 0;
  }
  size_t cap = (bytes + (bytes >> 2) + 255) & ~(size_t)255;  // 25% slack, 256-byte granules
  if (cudaMalloc(&s.p, cap) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&s.p, (bytes + 255) & ~(size_t)255) != cudaSuccess) {
      cudaGetLastError();
      s.p = nullptr;
      return nullptr;
    }
    cap = (bytes + 255) & ~(size_t)255;
  }
  s.cap = cap;
  return s.p;
}

void ed448_tables_free(DeviceCtx& dc);  // ed448_api.cu

}  // namespace capy

using namespace capy;

extern "C" {

int capy_version(void) { return 100; }

const char* capy_strerror(int status) {
  switch (status) {
    case CAPY_OK: return "ok";
    case CAPY_ERR_BAD_SECPARAM: return "unsupported security parameter (expected 224, 256, 384 or 512)";
    case CAPY_ERR_BAD_ARG: return "bad argument";
    case CAPY_ERR_CUDA: return "CUDA error (see capy_last_cuda_error)";
    case CAPY_ERR_BAD_POINT: return "input point is not on the curve";
    case CAPY_ERR_NO_DEVICE: return "no usable CUDA device";
    case CAPY_ERR_OOM: return "out of device memory";
    default: return "unknown status";
  }
}

int capy_gpu_init(const int* devices, int n_devices, capy_ctx** out_ctx) {
  if (!out_ctx) return CAPY_ERR_BAD_ARG;
  *out_ctx = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return CAPY_ERR_NO_DEVICE;
  }
  std::vector<int> devs;
  if (!devices || n_devices <= 0) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return CAPY_ERR_NO_DEVICE;
    devs.push_back(cur);
  } else {
    for (int i = 0; i < n_devices; i++) {
      if (devices[i] < 0 || devices[i] >= count) return CAPY_ERR_BAD_ARG;
      devs.push_back(devices[i]);
    }
  }
  capy_ctx* ctx = new (std::nothrow) capy_ctx();
  if (!ctx) return CAPY_ERR_OOM;
  ctx->devs.resize(devs.size());
  for (size_t i = 0; i < devs.size(); i++) {
    DeviceCtx& dc = ctx->devs[i];
    dc.dev = devs[i];
    DeviceGuard g(dc.dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dc.dev) != cudaSuccess) {
      capy_gpu_destroy(ctx);
      return CAPY_ERR_NO_DEVICE;
    }
    dc.sm_count = prop.multiProcessorCount;
    for (int s = 0; s < kNumStreams; s++) {
      if (cudaStreamCreateWithFlags(&dc.streams[s], cudaStreamNonBlocking) != cudaSuccess) {
        capy_gpu_destroy(ctx);
        return CAPY_ERR_CUDA;
      }
    }
  }
  *out_ctx = ctx;
  return CAPY_OK;
}

void capy_gpu_destroy(capy_ctx* ctx) {
  if (!ctx) return;
  for (DeviceCtx& dc : ctx->devs) {
    DeviceGuard g(dc.dev);
    cudaDeviceSynchronize();
    for (int s = 0; s < kNumStreams; s++)
      if (dc.streams[s]) cudaStreamDestroy(dc.streams[s]);
    for (Scratch& sc : dc.scratch)
      if (sc.p) cudaFree(sc.p);
    for (auto& kv : dc.prefix_cache) {
      if (kv.second.d_state) cudaFree(kv.second.d_state);
      if (kv.second.d_prefix) cudaFree(kv.second.d_prefix);
    }
    ed448_tables_free(dc);
  }
  delete ctx;
}

int capy_gpu_scrub(capy_ctx* ctx) {
  if (!ctx) return CAPY_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (DeviceCtx& dc : ctx->devs) {
    DeviceGuard g(dc.dev);
    CAPY_CUDA(ctx, cudaDeviceSynchronize());
    for (Scratch& sc : dc.scratch)
      if (sc.p) CAPY_CUDA(ctx, cudaMemset(sc.p, 0, sc.cap));
    CAPY_CUDA(ctx, cudaDeviceSynchronize());
  }
  return CAPY_OK;
}

int capy_gpu_device_count(const capy_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

const char* capy_last_cuda_error(const capy_ctx* ctx) { return ctx ? ctx->last_cuda_error.c_str() : ""; }

uint64_t capy_launch_count(const capy_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

void* capy_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void capy_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
