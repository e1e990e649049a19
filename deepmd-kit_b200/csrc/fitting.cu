// Operand splitting and recombination for the fitting net's GEMMs on the B200 tensor cores.
//
// The fitting net (deepmd/pt/model/network/mlp.py; D -> 240 -> 240 -> 240 -> 1 for water) is 2 MFLOP
// per atom of dense GEMM -- more than the table evaluation -- and fp64/fp32 parity (1e-10 / 1e-5)
// rules out a plain low-precision tensor-core product.  B200 runs FP64 GEMMs at ~33 TFLOP/s and
// FP32 (SIMT) at ~60, but INT8 at ~2000 TOP/s and TF32 at ~550 TFLOP/s (measured,
// gpurun_out/gemm_probe.log), so the products are made exact (or exact to 2^-22) by splitting
// the operands:
//
//   fp64  x = 2^Ex * sum_i X_i 2^(-7-8i),  X_i balanced base-256 digits (int8, -128..127)   [rows: per-row Ex]
//         w = 2^Ew * sum_j W_j 2^(-7-8j)                                                     [cols: per-col Ew]
//         x.w = 2^(Ex+Ew-14) * sum_d 2^(-8d) * sum_{i+j=d} X_i.W_j            (d < nslice)
//         Every X_i.W_j is an error-free int8 GEMM with int32 accumulation; the products of one order
//         d are ONE GEMM over the concatenated K axis [X_0|..|X_d] . [W_d;..;W_0].  The order sums
//         are recombined here in fp64 (Horner in 2^-8), with bias / tanh / resnet_dt fused in.
//   fp32  x = hi + lo with hi, lo representable in TF32:  x.w ~= hi.whi + lo.whi + hi.wlo
//         (3xTF32, relative error 2^-22), again one GEMM over a concatenated K axis.
//
// The GEMMs themselves are library calls (cuBLASLt through torch); these kernels are the glue that
// makes them usable at fp64/fp32 accuracy.  The first layer's left operand is produced already split
// by the tabulate forward's epilogue (tabulate.cu, desc_epilogue).
#include <cmath>

#include "common.cuh"

namespace dpb200 {
namespace {

// z = 2^(row_exp[r] + col_exp[c] - 14) * sum_d acc[d][r][c] 2^(-8d) + bias[c];
// a = tanh(z) (kept for the backward), y = a * idt (+ h).
__global__ void k_split_i8_combine(double* __restrict__ a_out, double* __restrict__ y_out,
                                   const int* __restrict__ acc, long long acc_stride, int ns,
                                   const int* __restrict__ row_exp, const int* __restrict__ col_exp,
                                   const double* __restrict__ bias, const double* __restrict__ idt,
                                   const double* __restrict__ h, long long n, int width, int act) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / width;
    const int c = (int)(e - r * width);
    double s = (double)acc[(long long)(ns - 1) * acc_stride + e];
    for (int d = ns - 2; d >= 0; --d) s = s * 0.00390625 + (double)acc[(long long)d * acc_stride + e];
    const int ex = row_exp[r] + col_exp[c] - 14;
    double z = ldexp(s, ex);
    if (bias) z += bias[c];
    if (act) {
      const double a = tanh(z);
      a_out[e] = a;
      double v = idt ? a * idt[c] : a;
      if (h) v += h[e];
      y_out[e] = v;
    } else {
      a_out[e] = z;
    }
  }
}

// Split an fp64 matrix [n][width] row-wise into `ns` balanced base-256 digit slices, most significant first:
// out[r][s][c] int8 (row stride ld_out bytes), row_exp[r].  One warp per row.
__global__ void k_split_i8_rows(signed char* __restrict__ out, long long ld_out, int* __restrict__ row_exp,
                                const double* __restrict__ x, long long ldx, long long n, int width, int ns) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long bias = 0;
  for (int k = 0; k < ns; ++k) bias = bias * 256ull + 128ull;
  const int P = 7 + 8 * (ns - 1);
  for (long long r = wid; r < n; r += nw) {
    const double* __restrict__ xr = x + r * ldx;
    double m = 0.;
    for (int c = lane; c < width; c += 32) m = fmax(m, fabs(xr[c]));
    int e = ((__double2hiint(m) >> 20) & 0x7ff) - 1023;
    e = __reduce_max_sync(kFull, e);
    int E = e + 2;
    E = E < -900 ? -900 : (E > 900 ? 900 : E);
    if (lane == 0) row_exp[r] = E;
    const double up = __hiloint2double((1023 + P - E) << 20, 0);
    signed char* __restrict__ o = out + r * ld_out;
    for (int c = lane; c < width; c += 32) {
      const unsigned long long q = (unsigned long long)__double2ll_rn(xr[c] * up) + bias;
      for (int s = 0; s < ns; ++s)
        o[(long long)s * width + c] = (signed char)((int)((q >> (8 * (ns - 1 - s))) & 255ull) - 128);
    }
  }
}

__device__ __forceinline__ float tf32r(float x) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// out[r] = [hi | lo | hi] (3*width floats, row stride ld_out), hi = tf32(x), lo = tf32(x - hi).
__global__ void k_split_tf32(float* __restrict__ out, long long ld_out, const float* __restrict__ x, long long ldx,
                             long long n, int width, int copies) {
  const long long tot = n * width;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / width;
    const int c = (int)(e - r * width);
    const float v = x[r * ldx + c];
    const float hi = tf32r(v);
    const float lo = tf32r(v - hi);
    float* __restrict__ o = out + r * ld_out;
    o[c] = hi;
    o[width + c] = lo;
    if (copies == 3) o[2 * width + c] = hi;
  }
}

}  // namespace
}  // namespace dpb200

extern "C" {

int dpb200_split_i8_combine_f64(double* a_out, double* y_out, const int* acc, long long acc_stride, int nslice,
                                const int* row_exp, const int* col_exp, const double* bias, const double* idt,
                                const double* h, long long nrow, int width, int activation,
                                dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nrow >= 0 && width >= 1 && nslice >= 1 && nslice <= 8, "split_i8_combine: bad shape");
  const long long n = nrow * width;
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(a_out && acc && row_exp && col_exp && (!activation || y_out), "split_i8_combine: null pointer");
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  k_split_i8_combine<<<grid, 256, 0, (cudaStream_t)stream>>>(a_out, y_out, acc, acc_stride, nslice, row_exp, col_exp,
                                                             bias, idt, h, n, width, activation);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

int dpb200_split_i8_rows_f64(signed char* out, long long ld_out, int* row_exp, const double* x, long long ldx,
                             long long nrow, int width, int nslice, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nrow >= 0 && width >= 1 && nslice >= 2 && nslice <= 8 && ld_out >= (long long)nslice * width &&
                  ldx >= width,
              "split_i8_rows: bad shape");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(out && row_exp && x, "split_i8_rows: null pointer");
  long long want = (nrow + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  k_split_i8_rows<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(out, ld_out, row_exp, x, ldx, nrow,
                                                                                   width, nslice);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

int dpb200_split_tf32_f32(float* out, long long ld_out, const float* x, long long ldx, long long nrow, int width,
                          int copies, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nrow >= 0 && width >= 1 && (copies == 2 || copies == 3) && ld_out >= (long long)copies * width &&
                  ldx >= width,
              "split_tf32: bad shape");
  const long long n = nrow * width;
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(out && x, "split_tf32: null pointer");
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  k_split_tf32<<<grid, 256, 0, (cudaStream_t)stream>>>(out, ld_out, x, ldx, nrow, width, copies);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // extern "C"
