// peaks.cu -- integer-pipe peak microbenchmarks for the roofline denominators.
//
// Standalone program:  peaks [iters]  -> one JSON object on stdout.
// Each kernel runs ILP independent dependency chains per thread of one instruction kind
// (inline PTX so ptxas cannot strength-reduce them), 256 threads x 8 blocks per SM, and
// reports warp-instructions issued per clock per SM (from clock64 on the device) and
// thread-ops per second (from CUDA events).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int ILP = 8;
constexpr int INNER = 16;  // unrolled repetitions of the ILP group per loop iteration

enum Kind { LOP3, SHF, IADD3, IMAD, IMADWIDE, IMADHI, MIX_LOP3_IMAD, MIX_LOP3_SHF, MIX_LOP3_IMADWIDE, MIX_KECCAK_FMA, DFMA, FFMA, MIX_DFMA_LOP3, MIX_DFMA_IMADWIDE, MIX_DFMA_IADD64, MIX_DFMA_DADD_IADD64, NKINDS };
static const char* kNames[NKINDS] = {"lop3", "shf", "iadd3", "imad", "imad_wide", "imad_hi", "mix_lop3_imad",
                                     "mix_lop3_shf", "mix_lop3_imadwide", "mix_2lop3_1imadwide", "dfma", "ffma", "mix_dfma_lop3",
                                     "mix_dfma_imadwide", "mix_dfma_iadd64", "mix_2dfma_dadd_2iadd64"};
// thread-ops per ILP-group element (mixes issue two instructions per element)
static const int kOpsPerElem[NKINDS] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 3, 1, 1, 2, 2, 2, 5};

template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(uint32_t* out, long long* cycles, int iters, uint32_t seed) {
  uint32_t a[ILP], a2[ILP], b = seed | 1u, c = seed * 2654435761u + 12345u;
  uint64_t w[ILP];
  double dd[ILP];
  float ff[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    a[i] = threadIdx.x * 977u + i * 131u + seed;
    w[i] = a[i];
    a2[i] = a[i] * 3u + 1u;
    dd[i] = (double)a[i];
    ff[i] = (float)a[i];
  }
  uint32_t bv[INNER];
#pragma unroll
  for (int r = 0; r < INNER; r++) bv[r] = seed * (2 * r + 3) + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (KIND == IMADWIDE || KIND == MIX_LOP3_IMADWIDE || KIND == MIX_KECCAK_FMA || KIND == MIX_DFMA_IMADWIDE) {
#pragma unroll
      for (int i = 0; i < ILP; i++) a2[i] ^= (uint32_t)(w[i] >> 7);  // keeps the products loop-variant
    }
#pragma unroll
    for (int r = 0; r < INNER; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (KIND == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (KIND == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
        if (KIND == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if (KIND == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (KIND == IMADWIDE) w[i] += (uint64_t)a2[i] * bv[r];  // schoolbook-shaped: ILP x INNER distinct products per iteration
        if (KIND == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (KIND == MIX_LOP3_IMAD) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2[i]) : "r"(b), "r"(c));
        }
        if (KIND == MIX_LOP3_SHF) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a2[i]) : "r"(b));
        }
        if (KIND == MIX_LOP3_IMADWIDE) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
          w[i] += (uint64_t)a2[i] * bv[r];  // schoolbook-shaped: ILP x INNER distinct products per iteration
        }
        if (KIND == MIX_KECCAK_FMA) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0xD2;" : "+r"(a[i]) : "r"(c), "r"(b));
          w[i] += (uint64_t)a2[i] * bv[r];  // schoolbook-shaped: ILP x INNER distinct products per iteration
        }
        if (KIND == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(1.0000001), "d"(0.5));
        if (KIND == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(ff[i]) : "f"(1.0000001f), "f"(0.5f));
        // FP64 pipe against the integer pipes: is DFMA (64 lanes/clk/SM on B200) a third pipe that overlaps LOP3 /
        // IMAD.WIDE / 64-bit adds?  (the question behind a double-precision-FMA field multiplier)
        if (KIND == MIX_DFMA_LOP3) {
          asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(1.0000001), "d"(0.5));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
        if (KIND == MIX_DFMA_IMADWIDE) {
          asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(1.0000001), "d"(0.5));
          w[i] += (uint64_t)a2[i] * bv[r];
        }
        if (KIND == MIX_DFMA_IADD64) {
          asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(1.0000001), "d"(0.5));
          asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)bv[r] << 20));
        }
        if (KIND == MIX_DFMA_DADD_IADD64) {  // the per-product work of the split-product scheme: 2 DFMA + 1 DADD + 2 x 64-bit add
          double hi, lo;
          asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(dd[i]), "d"(1.0000001), "d"(4503599627370496.0));
          asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(lo) : "d"(4503599627370497.0), "d"(hi));
          asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(lo) : "d"(dd[i]), "d"(1.0000001));
          w[i] += (uint64_t)__double_as_longlong(hi);
          w[(i + 1) % ILP] += (uint64_t)__double_as_longlong(lo);
        }
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) acc ^= a[i] ^ a2[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)dd[i] ^ (uint32_t)ff[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(int sms, int iters, uint32_t* d_out, long long* d_cyc, std::string& json) {
  const int blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  peak_kernel<KIND><<<blocks, threads>>>(d_out, d_cyc, iters / 8 + 1, 1u);  // warm-up
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  double cyc_mean = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    peak_kernel<KIND><<<blocks, threads>>>(d_out, d_cyc, iters, 7u + rep);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) {
      best_ms = ms;
      std::vector<long long> h(blocks);
      CK(cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
      double s = 0;
      for (long long v : h) s += (double)v;
      cyc_mean = s / blocks;
    }
  }
  const double elems_per_thread = (double)iters * INNER * ILP;
  const double thread_ops = elems_per_thread * kOpsPerElem[KIND] * (double)blocks * threads;
  // per SM: 8 resident blocks x 8 warps run concurrently for ~cyc_mean cycles
  const double warp_instr_per_sm = elems_per_thread * kOpsPerElem[KIND] * 8.0 * (threads / 32);
  char buf[512];
  snprintf(buf, sizeof buf,
           "\"%s\": {\"ms\": %.4f, \"thread_ops_per_s\": %.4e, \"warp_instr_per_clk_per_sm\": %.3f, "
           "\"lanes_per_clk_per_sm\": %.1f, \"block_cycles\": %.0f}",
           kNames[KIND], best_ms, thread_ops / (best_ms * 1e-3), warp_instr_per_sm / cyc_mean,
           32.0 * warp_instr_per_sm / cyc_mean, cyc_mean);
  if (json.size() > 1) json += ", ";
  json += buf;
  CK(cudaEventDestroy(e0));
  CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 2000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint32_t* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, (size_t)sms * 8 * 256 * 4));
  CK(cudaMalloc(&d_cyc, (size_t)sms * 8 * 8));
  std::string json = "{";
  run<LOP3>(sms, iters, d_out, d_cyc, json);
  run<SHF>(sms, iters, d_out, d_cyc, json);
  run<IADD3>(sms, iters, d_out, d_cyc, json);
  run<IMAD>(sms, iters, d_out, d_cyc, json);
  run<IMADWIDE>(sms, iters, d_out, d_cyc, json);
  run<IMADHI>(sms, iters, d_out, d_cyc, json);
  run<MIX_LOP3_IMAD>(sms, iters, d_out, d_cyc, json);
  run<MIX_LOP3_SHF>(sms, iters, d_out, d_cyc, json);
  run<MIX_LOP3_IMADWIDE>(sms, iters, d_out, d_cyc, json);
  run<MIX_KECCAK_FMA>(sms, iters, d_out, d_cyc, json);
  run<DFMA>(sms, iters, d_out, d_cyc, json);
  run<FFMA>(sms, iters, d_out, d_cyc, json);
  run<MIX_DFMA_LOP3>(sms, iters, d_out, d_cyc, json);
  run<MIX_DFMA_IMADWIDE>(sms, iters, d_out, d_cyc, json);
  run<MIX_DFMA_IADD64>(sms, iters, d_out, d_cyc, json);
  run<MIX_DFMA_DADD_IADD64>(sms, iters, d_out, d_cyc, json);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  char tail[256];
  snprintf(tail, sizeof tail, ", \"sm_count\": %d, \"device\": \"%s\", \"max_clock_mhz\": %.0f, \"iters\": %d}", sms,
           prop.name, clk_khz / 1000.0, iters);
  json += tail;
  printf("%s\n", json.c_str());
  return 0;
}
