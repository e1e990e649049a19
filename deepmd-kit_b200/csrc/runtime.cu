// Error state, device queries and ABI version for the dpb200 C ABI.
#include <mutex>

#include "common.cuh"

namespace dpb200 {
namespace {
thread_local std::string g_error;
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
}
