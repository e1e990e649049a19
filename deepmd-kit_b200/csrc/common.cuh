// Shared device/host helpers for the dpb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/dpb200.h"

namespace dpb200 {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define DPB_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return ::dpb200::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define DPB_REQUIRE(cond, msg)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::dpb200::set_error(msg);                \
      return DPB200_ERR_INVALID;               \
    }                                          \
  } while (0)

// Bookkeeping for dpb200_launch_count(): every host wrapper reports the kernels it enqueued.
void note_launches(int n);
void keep_async_pool();

// Number of SMs of the current device (cached per device).
int sm_count();

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Butterfly reduce-scatter: v[0..15] per lane -> lane l returns the warp-wide sum of v[(l >> 1) & 15].
template <typename FP>
__device__ __forceinline__ FP reduce_scatter16(FP (&v)[16], int lane) {
#pragma unroll
  for (int s = 16, n = 16; s >= 2; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int t = 0; t < n / 2; ++t) {
      const FP send = up ? v[t] : v[t + n / 2];
      const FP keep = up ? v[t + n / 2] : v[t];
      v[t] = keep + __shfl_xor_sync(kFull, send, s);
    }
  }
  return v[0] + __shfl_xor_sync(kFull, v[0], 1);
}
// Generic form: v[0..N-1] per lane (N = 2, 4, 8, 16, 32) -> lane l returns the warp-wide sum of
// v[l >> log2(32 / N)].
template <int N, typename FP>
__device__ __forceinline__ FP reduce_scatter(FP (&v)[N], int lane) {
  int s = 16;
#pragma unroll
  for (int n = N; n >= 2; n >>= 1, s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int t = 0; t < n / 2; ++t) {
      const FP send = up ? v[t] : v[t + n / 2];
      const FP keep = up ? v[t + n / 2] : v[t];
      v[t] = keep + __shfl_xor_sync(kFull, send, s);
    }
  }
  FP r = v[0];
  for (; s >= 1; s >>= 1) r += __shfl_xor_sync(kFull, r, s);
  return r;
}

// Streaming (evict-first) global stores for write-once outputs.
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(int* p, int v) { __stcs(p, v); }

template <typename FP>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
};
template <>
struct Vec4<double> {
  using type = double4;
};

__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace dpb200
