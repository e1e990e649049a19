// Error state, device queries and ABI version for the dpb200 C ABI.
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace dpb200 {
namespace {
thread_local std::string g_error;
std::atomic<long long> g_launches{0};

// Dependent-FMA chains, 8 independent accumulators per thread: peak FMA rate of the FP pipe.
template <typename FP>
__global__ void __launch_bounds__(256) k_fma_peak(FP* out, int iters, FP a, FP b) {
  FP x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = (FP)(threadIdx.x + q) * (FP)1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = x[q] * a + b;
    }
  }
  FP s = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += x[q];
  if (s == (FP)123456.789) out[0] = s;  // never true: keeps the chain alive
}

template <typename FP>
int fma_peak(double* tflops, cudaStream_t st) {
  FP* out = nullptr;
  DPB_CUDA(cudaMalloc((void**)&out, sizeof(FP)));
  const int blocks = sm_count() * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  DPB_CUDA(cudaEventCreate(&e0));
  DPB_CUDA(cudaEventCreate(&e1));
  double best = 0.;
  for (int rep = 0; rep < 4; ++rep) {
    DPB_CUDA(cudaEventRecord(e0, st));
    k_fma_peak<FP><<<blocks, threads, 0, st>>>(out, iters, (FP)0.999, (FP)1e-4);
    DPB_CUDA(cudaEventRecord(e1, st));
    DPB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    DPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  note_launches(4);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return DPB200_OK;
}
}  // namespace

void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Stream-ordered scratch (cudaMallocAsync) must not go back to the OS at every synchronisation:
// keep it in the device's default pool (set once per device).
void keep_async_pool() {
  static std::mutex mu;
  static bool done[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  std::lock_guard<std::mutex> lk(mu);
  if (done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
  done[dev] = true;
}

void set_error(const std::string& msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_error = std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what + " at " + file + ":" +
            std::to_string(line);
  // clear the sticky-free error so that the next call starts clean
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? DPB200_ERR_OOM : DPB200_ERR_CUDA;
}

int sm_count() {
  static std::mutex mu;
  static int cached[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (!have[dev]) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    cached[dev] = n;
    have[dev] = true;
  }
  return cached[dev];
}

}  // namespace dpb200

extern "C" {
const char* dpb200_last_error(void) { return dpb200::g_error.c_str(); }
int dpb200_abi_version(void) { return 1; }
long long dpb200_launch_count(void) { return dpb200::g_launches.load(std::memory_order_relaxed); }
void dpb200_count_replayed_launches(long long n) { dpb200::g_launches.fetch_add(n, std::memory_order_relaxed); }
int dpb200_fma_peak_f64(double* tflops, dpb200_stream_t stream) {
  DPB_REQUIRE(tflops != nullptr, "fma_peak: null output");
  return dpb200::fma_peak<double>(tflops, (cudaStream_t)stream);
}
int dpb200_fma_peak_f32(double* tflops, dpb200_stream_t stream) {
  DPB_REQUIRE(tflops != nullptr, "fma_peak: null output");
  return dpb200::fma_peak<float>(tflops, (cudaStream_t)stream);
}
}
