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
